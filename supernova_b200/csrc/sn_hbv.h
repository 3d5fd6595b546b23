// sn_hbv.h -- HyperBasevector construction from unipath edges (host side of a8/a9).
// Vertex discovery sorts 4E (K-1)-mer edge ends; numbering reproduces the reference's
// FIFO breadth-first order (paths/long/HBVFromEdges.cc:199-228), which is inherently
// sequential and tiny next to the k-mer stream, so it runs on the host between the
// device edge stage and the device pathing stage (DESIGN.md "HBV numbering").
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
    std::vector<std::vector<int32_t>> from, from_eo, to, to_eo;   // graph/Digraph.h adjacency, kept sorted
    std::vector<uint32_t> src;          // HBV edge -> unipath id << 1 | rc
    std::vector<int32_t> fwd, rev;      // unipath -> HBV edge (fwdEdgeXlat / revEdgeXlat)
    std::vector<int32_t> to_left, to_right;
    std::vector<int32_t> inv;           // HyperBasevector::Involution (paths/HyperBasevector.cc:685-697)
    // sequences of the HBV edges (edges_), fastb packing
    std::vector<uint8_t> epacked; std::vector<uint64_t> eoff; std::vector<uint32_t> elen;
};

void build_hbv(const Edges& edges, Hbv& out);

}  // namespace snh
