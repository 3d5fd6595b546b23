// sn_hbv.h -- HyperBasevector construction from unipath edges (host side of a8/a9).
// The numbering reproduces the reference's FIFO breadth-first order
// (paths/long/HBVFromEdges.cc:199-228); a FIFO traversal is sequential inside one connected
// component, so it runs on the host, all components in parallel, between the device stages that
// prepare it (vertex discovery, orders, components) and finish it (adjacency, involution).
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

// Records of the numbering loop.  GroupRec: one vertex (= distinct (K-1)-mer) with its edge ends
// in EEComp order (items = unipath << 1 | rc), one cache line.  ERec: one oriented unipath with
// its two vertices.  On the product path both are produced on the device (sn_hbvdev.cuh).
struct alignas(64) GroupRec { int32_t vid; uint32_t n; uint32_t items[8]; uint32_t pad[6]; };   // pad[0]: items already pushed
struct ERec { int32_t g1, g2, id; uint32_t pal; };
// One oriented unipath with everything the numbering loop needs on ONE cache line: its two
// vertices and the items (unipath << 1 | rc, EEComp order) of both.  info = n1 | n2 << 4 | pal << 8;
// a vertex with more than 6 items (n = 15) sends the loop to the vertex records instead.
struct alignas(64) ItemRec { int32_t g1, g2; uint32_t info; uint32_t it1[6], it2[6]; uint32_t pad; };
// connected components in the order the reference's outer loop discovers them
struct HbvComponents {
    uint64_t n_comp = 0;
    const uint32_t* start_item = nullptr;   // n_comp: unipath << 1 | rc
    const uint64_t* base_v = nullptr;       // n_comp + 1: first vertex id of each component
    const uint64_t* base_e = nullptr;       // n_comp + 1: first HBV edge id
};
// numbering (HBVBuilder::processQueue) of all components with `threads` host threads: fills
// n_vert, src, to_left, to_right, fwd, rev
// `layout` (or nullptr): the records are laid out along the graph (sn_hbvdev.cuh, k_lay_*): record layout[x] is
// item x, its lists hold record indices and its pad the item
// `part` of `n_parts` (multi-GPU): only every n_parts-th work unit is numbered here; out.src / to_left / to_right must come in
// zeroed, and what other parts number stays 0 (fwd / rev included), so the parts' arrays ADD UP to the complete ones
void number_hbv(const HbvComponents& comps, const ItemRec* items, const GroupRec* groups, uint64_t n_vertices, uint64_t n_unipaths, Hbv& out, unsigned threads,
                const uint32_t* layout = nullptr, unsigned part = 0, unsigned n_parts = 1, uint64_t min_items = 0);
// (min_items: components with fewer oriented edges are skipped -- the device numbered them -- and read as 0 like another part's)
#ifdef SN_HOSTSIM
// host-only construction of the whole HBV: exists in tests/hostsim only (the product library is built without it)
void build_hbv(const Edges& edges, Hbv& out);
#endif
// sequences of the HBV edges (edges_), fastb packing: epacked (padded), eoff[n+1], elen[n]
void hbv_edge_sequences(const Edges& edges, const Hbv& h, std::vector<uint8_t>& epacked, std::vector<uint64_t>& eoff, std::vector<uint32_t>& elen);

}  // namespace snh
