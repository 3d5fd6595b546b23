// sn_hbv.h -- HyperBasevector construction from unipath edges (host side of a8/a9).
// Vertex discovery groups the 4E (K-1)-mer edge ends with an open-addressing hash table;
// numbering reproduces the reference's FIFO breadth-first order
// (paths/long/HBVFromEdges.cc:199-228), which is inherently sequential and tiny next to the
// k-mer stream, so it runs on the host between the device edge stage and the device pathing
// stage (DESIGN.md "HBV numbering").
#pragma once
#include <cstdint>
#include <vector>

namespace snh {

struct Edges {                         // unipath edges, fastb packing, byte aligned per edge
    std::vector<uint8_t> packed;       // padded by 16 bytes
    std::vector<uint64_t> off;         // n+1
    std::vector<uint32_t> len;
    uint64_t n() const { return len.size(); }
};

struct Hbv {
    int32_t K = 48;
    int32_t n_vert = 0;
    // graph/Digraph.h from_/to_ with their edge objects, CSR, every list sorted by neighbour
    // vertex with insertion-order ties exactly as digraphE::AddEdge leaves them
    std::vector<uint32_t> from_start, to_start;             // n_vert+1
    std::vector<int32_t> from_v, from_e, to_v, to_e;        // n_hbv_edges each
    std::vector<uint32_t> src;          // HBV edge -> unipath id << 1 | rc
    std::vector<int32_t> fwd, rev;      // unipath -> HBV edge (fwdEdgeXlat / revEdgeXlat)
    std::vector<int32_t> to_left, to_right;
    std::vector<int32_t> inv;           // HyperBasevector::Involution (paths/HyperBasevector.cc:685-697)
};

// Vertex discovery precomputed on the device (k_hbv_keys / sort / k_hbv_mark): which group
// (= distinct (K-1)-mer) each edge end belongs to, the members of every group, a
// (length desc, first 32 bases) pre-order of the edges, and the palindrome flags.  When it is
// empty build_hbv derives the same things on the host.
struct HbvPre {
    std::vector<uint32_t> order;        // edge ids sorted by (len desc, first 32 bases); ties unresolved
    std::vector<uint8_t> pal;           // whole-edge canonical form == PALINDROME
    std::vector<int32_t> end_group;     // [4*e + 2*rc + distal] -> group or -1
    std::vector<uint32_t> group_start;  // n_groups+1
    std::vector<uint32_t> group_items;  // edge<<2 | rc<<1 | distal, grouped
    bool empty() const { return end_group.empty(); }
};

void build_hbv(const Edges& edges, const HbvPre& pre, Hbv& out);
inline void build_hbv(const Edges& edges, Hbv& out) { build_hbv(edges, HbvPre(), out); }
// sequences of the HBV edges (edges_), fastb packing: epacked (padded), eoff[n+1], elen[n]
void hbv_edge_sequences(const Edges& edges, const Hbv& h, std::vector<uint8_t>& epacked, std::vector<uint64_t>& eoff, std::vector<uint32_t>& elen);

}  // namespace snh
