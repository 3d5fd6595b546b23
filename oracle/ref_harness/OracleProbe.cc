// OracleProbe.cc -- TEST INFRASTRUCTURE (not product code).
//
// Reference-calling glue: a `main` that loads HEAD.fastb/.qualp/.bci exactly as
// the reference's DF driver does (10X/DF.cc:464-469 expands the barcode index to
// a per-read barcode ordinal) and calls the reference's own
// buildReadQGraph48 (paths/long/BuildReadQGraph48.cc:1688) with the argument
// values StageBuildGraph uses (10X/runstages/RunStages.cc:404-406).  It contains
// no re-implementation of the hot path; it is compiled against the UNMODIFIED
// reference sources under /root/reference by oracle/build_ref.sh into
// oracle/_ref/ (git-ignored).
//
// Usage: OracleProbe HEAD=<dir>/reads OUT=<dir> [PATHS=True|False] [INDEX=False|True] [DFSIDE=False|True] [MIN_QUAL=7]
//                    [MIN_FREQ=3] [MIN_BC=2] [IGN_BC_BELOW=0] [MSPEDGES=<edges.bv>]
//   DFSIDE=True (with INDEX and PATHS): also the other files DF leaves in a.48/ right after the hot path
//   (10X/DF.cc:573-600, 10X/WriteFiles.cc:16-60): a.k, a.hbx, a.fastb, a.kmers, a.pathsX, a.dup -- each
//   written by the reference's own code (HyperBasevectorX, ReadPathVecX, MarkDups).
//   MSPEDGES given: the reference's buildGraphFromMSP (BuildReadQGraph48.cc:1631-1684) instead -- the
//   production path, where the edges come from the tada stages -- on HEAD.fastb/.qualp and that edge file.
// The same source is linked twice by oracle/build_ref.sh: against the reference's BuildReadQGraph48.o
// (OracleProbe) and against integration/BuildReadQGraph48_b200.cc + libsupernova_b200.so (OracleProbe_b200,
// the drop-in test: same main, same reference closure, only that one translation unit replaced).
//   OUT/reads.fastb must exist (re-opened at BuildReadQGraph48.cc:1763).
// Outputs: OUT/a.hbv, OUT/tmp.paths (PATHS=True), OUT/kmers.kvec (when env
//   SN_KEEP_KVEC is set; needs the optional guard patch), OUT/stats/*.json,
//   OUT/fwd.xlat-free: the HBV itself carries the canonical edge numbering.
// Prints one line "ORACLE_SECONDS <wall seconds of the buildReadQGraph48 call>".
#include "MainTools.h"
#include "Basevector.h"
#include "feudal/ObjectManager.h"
#include "feudal/PQVec.h"
#include "paths/HyperBasevector.h"
#include "paths/long/ReadPath.h"
#include "paths/long/BuildReadQGraph48.h"
#include "10X/PathsIndex.h"
#include "10X/SecretOps.h"
#include "10X/DfTools.h"
#include "10X/paths/ReadPathVecX.h"
#include <chrono>

// Progress-dot printer declared in 10X/DfTools.h:466 (defined in DfTools.cc:617-635, a translation unit
// whose closure is most of DF): writePathsIndex only calls it when verbose.  An empty stand-in, so
// that DfTools.cc need not be linked; it computes nothing.
template <class T> void MakeDots(T& done, T& ndots, const T total) {}
template void MakeDots(int& done, int& ndots, const int total);
// The statistics sink MarkDups reports its percentages to (10X/DfTools.h:47; its one instance is defined in
// DfTools.cc:21, the same unlinked translation unit).  The instance, nothing more.
StatLogger StatLogger::gInst;

int main(int argc, char** argv)
{   RunTime();
    BeginCommandArguments;
    CommandArgument_String(HEAD);
    CommandArgument_String(OUT);
    CommandArgument_Bool_OrDefault(PATHS, True);
    CommandArgument_Int_OrDefault(MIN_QUAL, 7);
    CommandArgument_Int_OrDefault(MIN_FREQ, 3);
    CommandArgument_Int_OrDefault(MIN_BC, 2);
    CommandArgument_Bool_OrDefault(INDEX, False);
    CommandArgument_Bool_OrDefault(DFSIDE, False);
    CommandArgument_Int_OrDefault(IGN_BC_BELOW, 0);
    CommandArgument_String_OrDefault(MSPEDGES, "");
    EndCommandArguments;

    vecbvec reads(HEAD + ".fastb");
    ObjectManager<VecPQVec> quals(HEAD + ".qualp");
    vec<int64_t> bci;
    BinaryReader::readFile(HEAD + ".bci", &bci);
    vec<int32_t> bc(bci.back(), -1);
    for (int b = 0; b < bci.isize() - 1; b++)
        for (int64_t j = bci[b]; j < bci[b + 1]; j++) bc[j] = b;

    HyperBasevector hbv;
    ReadPathVec paths;
    auto t0 = std::chrono::steady_clock::now();
    if (MSPEDGES != "")
        buildGraphFromMSP(OUT, HEAD + ".fastb", HEAD + ".qualp", MSPEDGES, hbv, 48, paths);
    else
        buildReadQGraph48(OUT, "/reads", "", reads, quals, False, False,
                          MIN_QUAL, MIN_FREQ, IGN_BC_BELOW, MIN_BC, &bc, .75, 0, "",
                          True, False, &hbv, PATHS ? &paths : nullptr, 0.9, False);
    auto t1 = std::chrono::steady_clock::now();
    BinaryWriter::writeFile(OUT + "/a.hbv", hbv);
    if (INDEX && PATHS) {
        // what DF does next (10X/DF.cc:586-590): the involution and the edge -> reads index of the paths
        // (writePathsIndex, 10X/PathsIndex.cc:23-143) -- the reference's own code, called as DF calls it
        ReadPathVec rp(OUT + "/tmp.paths");
        vec<int> inv;
        hbv.Involution(inv);
        BinaryWriter::writeFile(OUT + "/a.inv", inv);
        vec<int> to_left, to_right;                                   // 10X/WriteFiles.cc:16-60 (a.to_left, a.to_right)
        hbv.ToLeft(to_left); hbv.ToRight(to_right);
        BinaryWriter::writeFile(OUT + "/a.to_left", to_left);
        BinaryWriter::writeFile(OUT + "/a.to_right", to_right);
        writePathsIndex(rp, inv, OUT, "a.paths.inv", "a.countsb", 15, false);
        if (DFSIDE) {
            Echo(ToString(hbv.K()), OUT + "/a.k");                     // WriteFiles.cc:33
            HyperBasevectorX hbx(hbv);                                  // DF.cc:576, WriteFiles.cc:40-41
            BinaryWriter::writeFile(OUT + "/a.hbx", hbx);
            {   vecbvec edges(hbv.Edges().begin(), hbv.Edges().end()); // WriteFiles.cc:46-47
                edges.WriteAll(OUT + "/a.fastb");   }
            {   vec<int> kmers(hbv.E());                               // WriteFiles.cc:48-51
                for (int e = 0; e < hbv.E(); e++) kmers[e] = hbv.Kmers(e);
                BinaryWriter::writeFile(OUT + "/a.kmers", kmers);   }
            ReadPathVecX pathsX;                                        // DF.cc:577-579 (InitializePathsXFromPaths, DfTools.cc:24-78,
            pathsX.append(rp, hbx);                                     //  appends the reads in order, piece by piece)
            pathsX.WriteAll(OUT + "/a.pathsX");                         // WriteFiles.cc:28
            vecbvec bases(HEAD + ".fastb");                             // DF.cc:594-600 (frag_reads_orig.fastb = the reads)
            VirtualMasterVec<PQVec> vmv(HEAD + ".qualp");
            vec<int32_t> bcd(bci.back(), -1);
            for (int b = 0; b < bci.isize() - 1; b++)
                for (int64_t j = bci[b]; j < bci[b + 1]; j++) bcd[j] = b;
            vec<Bool> dup; double interdup_rate;
            MarkDups(bases, vmv, pathsX, hbx, bcd, dup, interdup_rate);
            BinaryWriter::writeFile(OUT + "/a.dup", dup);
        }
    }
    std::cout << "ORACLE_SECONDS "
              << std::chrono::duration<double>(t1 - t0).count() << std::endl;
    return 0;
}
