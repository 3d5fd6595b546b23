// sn_formats.h -- host-side readers/writers for the on-disk formats the hot path
// must keep: feudal .fastb / .qualp / ReadPathVec, BINWRITE .bci,
// vec<basevector>, a.hbv.  New C++17 code; byte layouts verified against files
// written by the reference's own binaries (tests/test_formats.py).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace snf {

struct Fastb {                       // feudal/FeudalControlBlock.h:159-165, feudal/FieldVec.h:586-607
    std::vector<uint8_t> var;        // packed bases, 4 per byte, each read byte aligned
    std::vector<uint64_t> off;       // n+1 byte offsets into var
    std::vector<uint32_t> len;       // bases per read
};
struct Qualp {                       // feudal/PQVec.cc:87-127
    std::vector<uint8_t> var;        // PQVec blocks, 0-terminated per read
    std::vector<uint64_t> off;       // n+1
};

bool read_fastb(const std::string& path, Fastb& out, std::string& err);
bool read_qualp(const std::string& path, Qualp& out, std::string& err);
bool read_bci(const std::string& path, std::vector<int64_t>& bci, std::string& err);
bool write_fastb(const std::string& path, const Fastb& in, std::string& err);
bool write_qualp(const std::string& path, const Qualp& in, std::string& err);
bool write_bci(const std::string& path, const std::vector<int64_t>& bci, std::string& err);

// PQVecEncoder (feudal/PQVec.cc:17-127): appends the encoding of q[0..n) to out.
void pqvec_encode(const uint8_t* q, uint32_t n, std::vector<uint8_t>& out);
// decoder (feudal/PQVec.cc:129-187); returns the number of quals written
uint32_t pqvec_decode(const uint8_t* p, const uint8_t* pend, uint8_t* out, uint32_t cap);
// expand a .bci barcode index to the per-read ordinal DF passes down (10X/DF.cc:464-469)
bool expand_bci(const std::vector<int64_t>& bci, std::vector<int32_t>& bc);   // false: the index is not sorted / does not start at 0

// vec<basevector> (BINWRITE; feudal/BinaryStream.h:486-493, feudal/FieldVec.h:596-598)
bool write_bv(const std::string& path, const uint8_t* packed, const uint64_t* off, const uint32_t* len, uint64_t n, std::string& err);
bool read_bv(const std::string& path, Fastb& out, std::string& err);

// a.hbv (paths/HyperBasevector.cc:121-125, graph/DigraphTemplate.h:3091-3097)
// adjacency given in CSR form: from_start/to_start have n_vert+1 entries
bool write_hbv(const std::string& path, int32_t K, uint64_t n_vert, const uint32_t* from_start, const int32_t* from_v,
               const int32_t* from_e, const uint32_t* to_start, const int32_t* to_e,
               const uint8_t* packed, const uint64_t* off, const uint32_t* len, uint64_t n_edges, std::string& err);
// a.hbx (HyperBasevectorX: paths/HyperBasevector.cc:133-137, graph/Digraph.h:435-437, graph/DigraphTemplate.h:3107-3113):
// K, from_, to_, from_edge_obj_, to_edge_obj_ as MasterVec<SerfVec<int>>, edges_, to_left_, to_right_
bool write_hbx(const std::string& path, int32_t K, uint64_t n_vert, const uint32_t* from_start, const int32_t* from_v, const int32_t* from_e,
               const uint32_t* to_start, const int32_t* to_v, const int32_t* to_e,
               const uint8_t* packed, const uint64_t* off, const uint32_t* len, uint64_t n_edges,
               const int32_t* to_left, const int32_t* to_right, std::string& err);
// a.pathsX (ReadPathVecX::writeBinary, 10X/paths/ReadPathVecX.cc:976-996): skip = 10, start_rid = 0, next_start_rid,
// the two sizes, ZipIndex (i64), ZippedData
bool write_pathsx(const std::string& path, uint64_t n_reads, const int64_t* index, uint64_t n_index, const uint8_t* data, uint64_t n_bytes, std::string& err);
// vec<unsigned char> (a.dup: vec<Bool>) and a plain text file (a.k)
bool write_vec_u8(const std::string& path, const uint8_t* v, uint64_t n, std::string& err);
bool write_text(const std::string& path, const std::string& text, std::string& err);
// feudal ReadPathVec (paths/long/ReadPath.h:61-63, feudal/FeudalFileWriter.cc:100-121)
bool write_paths(const std::string& path, uint64_t n, const int32_t* offset, const uint64_t* poff, const int32_t* edges, std::string& err);
// whole file into memory; a gzip file (.gz) is inflated with zlib (the reference pipes it through zcat)
bool read_text_maybe_gz(const std::string& path, std::vector<char>& out, std::string& err);
// VecULongVec (a.paths.inv: feudal, FCB sizeofFixed 0, sizeofX 16, sizeofA 8; 10X/PathsIndex.cc:75,107) and
// vec<vec<int>> with one inner vector (a.countsb, :133)
bool write_ulongvecs(const std::string& path, uint64_t n, const uint64_t* ids, const uint64_t* off /*n+1, in elements*/, std::string& err);
bool write_vec_vec_int(const std::string& path, const std::vector<int32_t>& v, std::string& err);
// vec<int> (a.inv)
bool write_vec_int(const std::string& path, const std::vector<int32_t>& v, std::string& err);

}  // namespace snf
