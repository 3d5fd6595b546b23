// BuildReadQGraph48_b200.cc -- the reference-side binding of libsupernova_b200.so.
//
// A maintainer drops this file into lib/assembly/src/paths/long/ IN PLACE OF BuildReadQGraph48.cc and
// links -lsupernova_b200: it defines the two functions DF's StageBuildGraph calls
// (10X/runstages/RunStages.cc:398-413) with the reference's own signatures
// (paths/long/BuildReadQGraph48.h:24-40) and does their work through the C ABI of
// include/supernova_b200.h.  Nothing else in DF changes.
//
// It is compiled against the UNMODIFIED reference headers by oracle/build_ref.sh (into the test harness
// oracle/_ref/OracleProbe_b200 = the reference closure minus BuildReadQGraph48.o plus this file) and run
// by tests/test_gpu_shim.py, which diffs its a.hbv / tmp.paths against the stock reference binary.
//
// Only accessors that exist in the reference are used: the reads and quals travel as the files DF
// loaded them from (FirstLoadData, 10X/DF.cc:313-316; the original re-opens the same files for its own
// pathing stage, BuildReadQGraph48.cc:1760-1763), the barcodes as the vec<int32_t> the caller holds, and
// the HyperBasevector comes back through the reference's own BinaryReader.
#include "MainTools.h"
#include "Basevector.h"
#include "feudal/BinaryStream.h"
#include "feudal/ObjectManager.h"
#include "feudal/PQVec.h"
#include "paths/HyperBasevector.h"
#include "paths/long/ReadPath.h"
#include "paths/long/BuildReadQGraph48.h"
#include "system/System.h"
#include "supernova_b200.h"

namespace {

struct Ctx {
    sn_ctx* c = nullptr;
    Ctx() { if (sn_ctx_create(&c, 0)) FatalErr(String("supernova_b200: ") + sn_last_error(nullptr)); }
    ~Ctx() { sn_ctx_destroy(c); }
    void ok(int rc) { if (rc) FatalErr(String("supernova_b200: ") + sn_last_error(c)); }   // FatalErr -> exit(1), as the original
};

void hbvFromFile(Ctx& x, String const& work_dir, HyperBasevector* pHBV)
{
    String tmp = work_dir + "/b200.tmp.hbv";
    x.ok(sn_write_hbv(x.c, tmp.c_str()));
    BinaryReader::readFile(tmp, pHBV);                   // the reference's own reader rebuilds to_ etc.
    Remove(tmp);
}

}  // namespace

void buildReadQGraph48( String const& work_dir, String const& read_head, std::string const mspFilename,
        vecbvec& reads, ObjectManager<VecPQVec>& quals, bool doFillGaps, bool doJoinOverlaps,
        unsigned minQual, unsigned minFreq, int64_t const ignBcBelow, unsigned minBC,
        vec<int32_t> const* bcp, double minFreq2Fract, unsigned maxGapSize, String const& refFasta,
        bool useNewAligner, bool repathUnpathed, HyperBasevector* pHBV, ReadPathVec* pPaths,
        float const memFrac, bool const VERBOSE )
{
    ForceAssertEq(doFillGaps, False);                    // BuildReadQGraph48.cc:1703-1705
    ForceAssertEq(doJoinOverlaps, False);
    ForceAssertEq(repathUnpathed, False);
    ForceAssert(useNewAligner);
    ForceAssert(refFasta.empty());
    if (mspFilename.size() > 0) FatalErr("old Msp not supported");          // :1710
    cout << Date() << ": loading reads (supernova_b200)." << endl;

    Ctx x;
    String fastb = work_dir + read_head + ".fastb";      // the file the original re-opens at :1763
    String qualp = quals.filename();                     // ... and at :1760
    ForceAssert(!bcp || bcp->size() == reads.size());
    x.ok(sn_load_read_files_bc(x.c, fastb.c_str(), qualp.c_str(), bcp ? &(*bcp)[0] : nullptr, bcp ? bcp->size() : 0));
    sn_params prm; prm.min_qual = minQual; prm.min_freq = minFreq; prm.min_bc = minBC; prm.ign_bc_below = ignBcBelow;
    x.ok(sn_count_kmers(x.c, &prm));                     // createDict up to the KmerVec (:218-292)
    Mkdir777(work_dir + "/stats");
    x.ok(sn_write_kmer_spectrum(x.c, (work_dir + "/stats/histogram_kmer_count.json").c_str()));   // :199-216
    sn_counts cnt; x.ok(sn_get_counts(x.c, &cnt));
    cout << Date() << ": dictionary covers " << ToStringAddCommas(cnt.n_kmers) << " kmers" << endl;
    if (cnt.n_kmers == 0) FatalErr("no valid k-mers");    // the original dies inside createDict on an empty KmerVec
    x.ok(sn_build_edges(x.c));                           // recomputeAdjacencies + buildEdges (:320-321,514-541)
    x.ok(sn_build_hbv(x.c));                             // buildHBVFromEdges (HBVFromEdges.cc:244-296)
    if (pPaths) {
        Destroy(reads);                                  // ownership as in :1751
        quals.unload();                                  // :1757
        x.ok(sn_path_reads(x.c));                        // pathReads (:1440-1469)
        x.ok(sn_write_paths(x.c, (work_dir + "/tmp.paths").c_str()));     // paths stay ON DISK (DF.cc:579-584)
    }
    hbvFromFile(x, work_dir, pHBV);
}

void buildGraphFromMSP( String const& work_dir, String const& reads_name, String const& quals_name,
        String const& MSPEDGES, HyperBasevector& hbv, const int K, ReadPathVec& paths )
{
    ForceAssertEq(K, 48);
    cout << Date() << ": reading MSP edge file " << MSPEDGES << " (supernova_b200)" << endl;
    Ctx x;
    x.ok(sn_load_read_files_bc(x.c, reads_name.c_str(), quals_name.c_str(), nullptr, 0));
    x.ok(sn_build_graph_from_edges(x.c, MSPEDGES.c_str()));              // mspEdgesToHBV + dictionary of the edge k-mers (:1647-1664)
    x.ok(sn_path_reads(x.c));                                            // :1680
    x.ok(sn_write_paths(x.c, (work_dir + "/tmp.paths").c_str()));
    hbvFromFile(x, work_dir, &hbv);
}
