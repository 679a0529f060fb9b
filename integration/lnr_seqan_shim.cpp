// lnr_seqan_shim.cpp -- the reference-side shim of INTEGRATION.md, as a compiled file.
//
// It gives the UNMODIFIED reference (xp3i4/linear) its GPU path with zero source changes: the two functions below have
// the reference's own signatures, and the reference's objects are linked with their definitions of these two symbols
// weakened (objcopy --weaken-symbol, oracle/build_ref.sh), so every call site of the reference -- Mapper::createIndex
// (mapper.cpp:325), Mapper::p_calRecords (mapper.cpp:447), map_ (mapper.cpp:849) -- lands here:
//
//   createIndexDynamic(genomes, index, gstr, gend, threads, efficient)   index_util.cpp:2478
//       -> lnr_genome_upload + lnr_index_build (+ lnr_features_build on first use); no host index is built at all
//   apxMap(index, read, anchors, hit, f1, f2, apx_gaps, cords_str, cords_end, cords_info, f_chain, pm_g, pm_pmp)
//       pmpfinder.cpp:2709 -> lnr_apxmap_batch on the calling thread's context; cords_str / cords_end are filled exactly as
//       the reference fills them, so mapGaps, reformCords, cords2BamLink and the APF / SAM writers consume them unchanged.
//
// The reference's host stages still get what they read besides the cords: the genome features f2 (createFeatures in
// linear.cpp:14 stays the reference's own host code) and the read features f1 (p_calRecords computes them itself).
// One call per read keeps the shim free of any restated reference logic; a production integration batches a whole block
// of reads per call -- same entry point, same results.
//
// This file is OUR code. It includes the reference's headers where they lie (never copied) and is compiled by
// oracle/build_ref.sh next to the reference objects, only where the reference tree is present.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "base.h"
#include "cords.h"
#include "index_util.h"
#include "pmpfinder.h"

#include "lnr_b200.h"

using namespace seqan;

namespace {

struct GpuPath
{
    std::mutex mu;
    int device = 0;
    lnr_ctx * build_ctx = nullptr;
    lnr_genome * genome = nullptr;
    lnr_index * index = nullptr;
    lnr_feats * feats = nullptr;
    int feats_type = -1;
    unsigned threads_sem = 1;
    std::vector<lnr_ctx *> pool;        // contexts handed to the calling threads (one per thread, created on first use)
};
GpuPath G;

[[noreturn]] void die(const char * what, int rc, lnr_ctx * ctx)
{
    // there is no CPU fallback: a failing GPU call ends the program loudly
    std::fprintf(stderr, "lnr_b200 shim: %s failed (%d): %s\n", what, rc, ctx ? lnr_last_error(ctx) : "");
    std::abort();
}

lnr_ctx * thread_ctx()
{
    thread_local lnr_ctx * ctx = nullptr;
    if (!ctx)
    {
        int rc = lnr_ctx_create(G.device, &ctx);
        if (rc) die("lnr_ctx_create", rc, nullptr);
        std::lock_guard<std::mutex> lk(G.mu);
        G.pool.push_back(ctx);
    }
    return ctx;
}

void ensure_features(int fs_type)
{
    std::lock_guard<std::mutex> lk(G.mu);
    if (G.feats && G.feats_type == fs_type) return;
    if (G.feats) lnr_features_destroy(G.feats);
    // FeaturesDynamic::fs_type: typeFeatures1_32 = 1, typeFeatures2_48 = 2 (pmpfinder.cpp:31-33) = the -f values
    int rc = lnr_features_build(G.build_ctx, G.genome, fs_type, G.threads_sem, &G.feats);
    if (rc) die("lnr_features_build", rc, G.build_ctx);
    G.feats_type = fs_type;
}

}  // namespace

// index_util.cpp:2478
bool createIndexDynamic(StringSet<String<Dna5> > & seqs, IndexDynamic & index, unsigned gstr, unsigned gend, unsigned threads, bool efficient)
{
    (void)efficient;
    if (const char * e = std::getenv("LNR_DEVICE")) G.device = std::atoi(e);
    int rc = lnr_ctx_create(G.device, &G.build_ctx);
    if (rc) die("lnr_ctx_create", rc, nullptr);
    if (gstr != 0 || gend != length(seqs)) die("createIndexDynamic over a sub-range of the genome", LNR_E_UNSUPPORTED, nullptr);
    std::vector<const uint8_t *> ptr;
    std::vector<uint64_t> len;
    for (unsigned i = gstr; i < gend; i++)
    {
        ptr.push_back((const uint8_t *)&seqs[i][0]);        // String<Dna5>: one byte per base, ordinals 0..4 (base.h:106)
        len.push_back(length(seqs[i]));
    }
    if ((rc = lnr_genome_upload(G.build_ctx, (uint32_t)ptr.size(), ptr.data(), len.data(), &G.genome))) die("lnr_genome_upload", rc, G.build_ctx);
    G.threads_sem = threads ? threads : 1;
    if ((rc = lnr_index_build(G.build_ctx, G.genome, index.isHIndex() ? 2 : 1, G.threads_sem, &G.index))) die("lnr_index_build", rc, G.build_ctx);
    return true;
}

// pmpfinder.cpp:2709
uint64_t apxMap(IndexDynamic & index, String<Dna5> & read, Anchors & anchors, String<uint64_t> & hit, StringSet<FeaturesDynamic> & f1,
                StringSet<FeaturesDynamic> & f2, String<UPair> & apx_gaps, String<uint64_t> & cords_str, String<uint64_t> & cords_end,
                String<CordInfo> & cords_info, int f_chain, GlobalParms & pm_g, PMPParms & pm_pmp)
{
    (void)index; (void)anchors; (void)hit; (void)f1; (void)cords_info; (void)pm_g;
    if (!G.index) die("apxMap before createIndexDynamic", LNR_E_ARG, nullptr);
    const int fs_type = f2[0].fs_type;
    if (!G.feats || G.feats_type != fs_type) ensure_features(fs_type);
    lnr_ctx * ctx = thread_ctx();
    clear(apx_gaps);                                         // pmpfinder.cpp:2731; mapGaps recomputes them (gap.cpp:444)
    const uint64_t L = length(read);
    uint64_t off[2] = {0, L};
    std::vector<uint64_t> cords(L / 4 + 64);
    uint64_t coff[2] = {0, 0};
    lnr_params prm;
    std::memset(&prm, 0, sizeof prm);
    prm.preset = pm_pmp.pm_cah.thd_stop_chain_len_ratio > 0.0f ? 0 : 1;   // mapper.cpp:181-195: -p 0 => 0.7, -p 1/2 => 0
    prm.feature_type = fs_type;
    if (!f_chain)
    {
        // -c 0 (alg_type 1, pmpfinder.cpp:2773-2787). This mode reads AND changes the caller's per-thread PMPParms:
        // GetDHitListParms is (20, 1) as constructed and (10, 999) once any read of the thread has needed the second attempt
        // (toggle(1) ... toggle(0), :2782-2784). The library takes the state as a parameter; the shim keeps the object in step.
        prm.no_chain = 1;
        prm.gdl_state = pm_pmp.pm_gdl.thd_list_n == 10 ? 1 : 0;
    }
    int rc = lnr_apxmap_batch(ctx, G.index, G.feats, &prm, 1, (const uint8_t *)&read[0], off, cords.data(), coff, cords.size(), nullptr);
    if (rc) die("lnr_apxmap_batch", rc, ctx);
    if (!f_chain)
    {
        uint64_t counters[8];
        if (lnr_last_batch_counters(ctx, counters) == 0 && counters[7] > 0) { pm_pmp.toggle(1); pm_pmp.toggle(0); }   // the second attempt ran
    }
    const uint64_t n = coff[1];
    const uint64_t w = fs_type == 1 ? 192 : 96;              // getFeatureWindowSize(f1), pmpfinder.cpp:2724
    const uint64_t d = (w << 20) | w;                        // shift_cord(0, w, w), pmpfinder.cpp:2790
    // the reference appends to cords_str (initCords on an empty string); callers pass it empty (mapper.cpp:419-421)
    resize(cords_str, n);
    resize(cords_end, n);
    for (uint64_t i = 0; i < n; i++) { cords_str[i] = cords[i]; cords_end[i] = cords[i] + d; }
    return 0;
}
