// CPU-side logic check of the PRODUCT code (not the oracle): compiles linear_b200/csrc/lnr_*.h as plain
// C++ with a single-lane "warp" and drives the same per-element functions the CUDA kernels call
// (idx_sample, feat_cell, seed_sample, phase_map / phase_mid / phase_finish) with sequential loops in
// place of the kernel grids. Used by the `-m "not gpu"` tests to compare the host logic of the pipeline
// with the oracle on a machine without a GPU. Nothing here ships; the kernels are in lnr_kernels.cu.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../linear_b200/csrc/lnr_core.h"
#include "../../linear_b200/csrc/lnr_pipeline.h"
#include "../../linear_b200/csrc/lnr_bamrec.h"

using namespace lnr;

namespace {

struct Emu
{
    std::vector<std::vector<u8> > g;
    std::vector<u64> glen;
    unsigned T;
    float stop_ratio;
    std::vector<i32> dir;
    std::vector<u64> hs;
    std::vector<std::vector<F96> > f2;
    std::vector<const F96 *> f2p;
    int ft = 2;                              // feature type: 2 = 2_48, 1 = 1_32
    std::vector<std::vector<i16> > s2;
    std::vector<const i16 *> s2p;
    std::vector<u32> nf2;
    std::vector<u64> out;
    std::vector<i32> outf;
    std::vector<u32> bins;
    std::vector<int64_t> recs;
    std::vector<u64> cigs;
};

struct SeqAcc
{
    const u8 * s; i64 len;
    int operator()(i64 p) const { return (p >= 0 && p < len) ? s[p] : 0; }
};
struct RcAcc   // reverse complement view (_compltRvseStr base.cpp:335)
{
    const u8 * s; i64 len;
    int operator()(i64 p) const
    {
        if (p < 0 || p >= len) return 0;
        int c = s[len - 1 - p];
        return c < 4 ? 3 - c : 4;
    }
};

template <class Acc> void build_feats(Acc acc, u32 n_entries, std::vector<F96> & f)
{
    f.resize(n_entries);
    std::vector<u64> lo(n_entries + 2);
    std::vector<u32> hi(n_entries + 2);
    for (u32 c = 0; c < n_entries + 2; c++) feat_cell(acc, 16 * (i64)c, lo[c], hi[c]);
    for (u32 e = 0; e < n_entries; e++) f[e] = feat_entry(lo[e] + lo[e + 1] + lo[e + 2], hi[e] + hi[e + 1] + hi[e + 2]);
}

void build_dindex(Emu & E)
{
    std::vector<u32> cnt(kDirSize, 0);
    std::vector<std::pair<u32, u64> > recs;
    for (size_t ci = 0; ci < E.g.size(); ci++)
    {
        SeqAcc acc = {E.g[ci].data(), (i64)E.glen[ci]};
        for (unsigned c = 0; c < E.T; c++)
        {
            IdxChunk ch;
            ch.base_off = 0; ch.len = (i64)E.glen[ci]; ch.contig = (u32)ci; ch.sample0 = 0;
            idx_chunk_range(ch.len, E.T, c, ch.t_str, ch.n_samples);
            if (ch.n_samples <= 0) continue;
            ch.kskip = hash_init_skip<kSpanD>(acc, ch.t_str, ch.len);
            ch.bias = ch.kskip ? selector_bias<kSpanD>(acc, ch.t_str + ch.kskip, ch.t_str) : 0;
            // emit(m) = (m - run_start(m)) even, run = maximal stretch of equal X
            u32 prevX = 0; i64 run_start = 0;
            for (i64 m = 0; m < ch.n_samples; m++)
            {
                u32 X; u64 rec;
                idx_sample(acc, ch, m, X, rec);
                if (m == 0 || X != prevX) run_start = m;
                prevX = X;
                if (((m - run_start) & 1) == 0) { cnt[X]++; recs.push_back(std::make_pair(X, rec)); }
            }
        }
    }
    E.dir.assign(kDirSize, 0);
    i64 sum = 0;
    for (u32 i = 0; i < kDirSize; i++)
    {
        u32 c = cnt[i] > (u32)kIdxOmit ? 0 : cnt[i];
        E.dir[i] = (i32)sum;
        sum += c;
        cnt[i] = c;
    }
    E.hs.assign(sum, 0);
    std::vector<i32> fill(E.dir);
    for (size_t i = 0; i < recs.size(); i++)
        if (cnt[recs[i].first]) E.hs[fill[recs[i].first]++] = recs[i].second;
    for (u32 i = 0; i + 1 < kDirSize; i++)
        if (E.dir[i + 1] - E.dir[i] > 1) std::sort(E.hs.begin() + E.dir[i], E.hs.begin() + E.dir[i + 1]);
}

SeedTask make_task(const SeqAcc & acc, u32 read, u32 str, u32 end, u32 alpha)
{
    SeedTask t;
    t.read = read; t.str = str; t.end = end; t.alpha = alpha; t.pad = 0; t.sample0 = 0;
    t.n_samples = seed_task_samples(str, end, alpha);
    t.kskip = (u32)hash_init_skip<kSpanD>(acc, 0, acc.len);
    t.bias = selector_bias<kSpanD>(acc, t.kskip, (i64)str + kSpanD);
    return t;
}

// seed_count + seed_fill kernels, sequentially; anchors[0] is the sentinel slot
void seed(const Emu & E, const SeqAcc & acc, const SeedTask & t, std::vector<u64> & anchors)
{
    anchors.assign(1, 0);
    u32 xprev = 0;
    for (u32 m = 1; m <= t.n_samples; m++)
    {
        SeedVal sv; u32 k;
        seed_sample(acc, t, m, sv, k);
        bool skip = sv.X == xprev;
        xprev = sv.X;
        if (skip) continue;
        for (i32 i = E.dir[sv.X]; i < E.dir[sv.X + 1]; i++)
            if (ykey_match((u32)(E.hs[i] & kMaskY), sv.Y)) anchors.push_back(val2anchor(E.hs[i], k, (u64)acc.len, sv.strand));
    }
}

struct ReadRun
{
    std::vector<u8> arena, a2;
    std::vector<u64> A, B, cords, dbg_hits;
    std::vector<F96> f1[2];
    std::vector<i16> s1[2];
    u32 hist[lnr::kWarpSmemWords] = {0};
};
template <class Acc> void build_feats32(const Acc & acc, u32 n, u32 written, std::vector<i16> & f)
{
    f.assign(n, 0);
    for (u32 i = 0; i < written && i < n; i++) f[i] = feat32_entry(acc, 16 * (i64)i);
}

// the per-read orchestration of kernels A and B (apxMap, pmpfinder.cpp:2709)
int run_read(Emu & E, const u8 * read, u64 L, std::vector<u64> & cords_out, std::vector<u64> * hits_out, int stop_after_first)
{
    Warp w = {0, 1u};
    SeqAcc acc = {read, (i64)L};
    RcAcc rc = {read, (i64)L};
    ReadRun R;
    PipeIn in;
    in.read = read; in.L = (u32)L;
    in.ft = E.ft; in.win = E.ft == 1 ? (u32)kWin32 : (u32)kWin;
    in.f1[0] = in.f1[1] = nullptr; in.s1[0] = in.s1[1] = nullptr; in.f2 = nullptr; in.s2 = nullptr;
    if (E.ft == 1)
    {
        u32 nf = feat32_count(L);
        build_feats32(acc, nf, feat32_written_serial(L), R.s1[0]);
        build_feats32(rc, nf, feat32_written_serial(L), R.s1[1]);
        in.s1[0] = R.s1[0].data(); in.s1[1] = R.s1[1].data(); in.nf1 = nf;
        in.s2 = E.s2p.data();
    }
    else
    {
        u32 nf = feat_count_read(L);
        build_feats(acc, nf, R.f1[0]);
        build_feats(rc, nf, R.f1[1]);
        in.f1[0] = R.f1[0].data(); in.f1[1] = R.f1[1].data(); in.nf1 = nf;
        in.f2 = E.f2p.data();
    }
    in.nf2 = E.nf2.data(); in.stop_ratio = E.stop_ratio;
    R.arena.resize(64 << 20);
    Arena ar = {R.arena.data(), R.arena.size(), 0, 0};
    int cap = 16 + (int)(L / 4);
    R.cords.assign(cap, 0);
    int nc = 0;
    PipeCounters cnt = {0, 0};
    if (E.bins.empty()) E.bins.assign(kNumBins, 0);
    SeedTask t = make_task(acc, 0, 0, (u32)L, 15);
    seed(E, acc, t, R.A);
    R.B.assign(R.A.size(), 0);
    std::vector<u64> dbg(R.A.size() + 2);
    u32 ndbg = 0;
    // stage 1 (k_map_hits) hands the hits over in A, stage 2 (k_map_extend) runs the window extension
    u32 nh = 0;
    if (phase_map(w, ar, R.hist, E.bins.data(), in, R.A.data(), R.B.data(), (int)R.A.size(), 0, L & kMaskY, 0,
                  R.cords.data(), nc, cap, dbg.data(), &ndbg, (u32)dbg.size(), cnt, R.A.data(), &nh))
        return 1;
    if (!path_dst_2(w, in, R.A.data(), (int)nh, R.cords.data(), nc, cap, 0, L & kMaskY, cnt)) return 1;
    if (hits_out) hits_out->assign(dbg.begin(), dbg.begin() + ndbg);
    if (stop_after_first) { cords_out.assign(R.cords.begin(), R.cords.begin() + nc); return 0; }
    arena_reset(ar);
    YPair * str_ends = arena_alloc<YPair>(ar, nc + 2);
    Blk * sep = arena_alloc<Blk>(ar, nc + 2);
    YPair * gaps = arena_alloc<YPair>(ar, L / 1000 + 4);
    int n_sep = 0, n_gaps = 0;
    int remap = phase_mid_w(w, L, R.cords.data(), nc, str_ends, sep, n_sep, gaps, n_gaps, (int)(L / 1000 + 4), in.win);
    if (remap < 0) return 1;
    std::vector<YPair> gv(gaps, gaps + n_gaps);
    std::vector<Blk> sepv(sep, sep + n_sep);
    if (remap)
    {
        for (int i = 0; i < n_gaps; i++)
        {
            SeedTask t2 = make_task(acc, 0, (u32)(gv[i].first & kMaskY), (u32)gv[i].second, 7);
            seed(E, acc, t2, R.A);
            R.B.assign(R.A.size(), 0);
            if (phase_map(w, ar, R.hist, E.bins.data(), in, R.A.data(), R.B.data(), (int)R.A.size(), gv[i].first & kMaskY,
                          gv[i].second & kMaskY, 1, R.cords.data(), nc, cap, 0, 0, 0, cnt, R.A.data(), &nh))
                return 1;
            if (!path_dst_2(w, in, R.A.data(), (int)nh, R.cords.data(), nc, cap, gv[i].first & kMaskY, gv[i].second & kMaskY, cnt)) return 1;
        }
        sepv.assign(nc + 2, Blk());
        int dummy = 0;
        n_sep = gather_blocks_w(w, R.cords.data(), nc, (YPair *)0, dummy, sepv.data(), L, 1000, in.win, 1);
    }
    arena_reset(ar);
    Blk * sp1 = arena_alloc<Blk>(ar, n_sep + 1);
    Blk * sp2 = arena_alloc<Blk>(ar, n_sep + 1);
    i32 * sc1 = arena_alloc<i32>(ar, n_sep + 1);
    i32 * sc2 = arena_alloc<i32>(ar, n_sep + 1);
    BlockScratch s1, s2;
    block_scratch_alloc(ar, s1, n_sep + 1);
    block_scratch_alloc(ar, s2, n_sep + 1);
    u64 * tmp = arena_alloc<u64>(ar, cap);
    if (ar.failed) return 1;
    for (int i = 0; i < n_sep; i++) sp1[i] = sepv[i];
    phase_finish_w(w, L, R.cords.data(), nc, sp1, n_sep, sp2, sc1, sc2, s1, s2, tmp, in.win);
    cords_out.assign(R.cords.begin(), R.cords.begin() + nc);
    return 0;
}

// -c 0 (apxMap with f_chain = 0): the product's phase_c0 / c0_finish driven the way apxmap_core drives k_map_c0
int run_read_c0(Emu & E, const u8 * read, u64 L, std::vector<u64> & cords_out, int gdl_state)
{
    Warp w = {0, 1u};
    SeqAcc acc = {read, (i64)L};
    RcAcc rc = {read, (i64)L};
    ReadRun R;
    PipeIn in;
    in.read = read; in.L = (u32)L;
    in.ft = E.ft; in.win = E.ft == 1 ? (u32)kWin32 : (u32)kWin;
    in.f1[0] = in.f1[1] = nullptr; in.s1[0] = in.s1[1] = nullptr; in.f2 = nullptr; in.s2 = nullptr;
    if (E.ft == 1)
    {
        u32 nf = feat32_count(L);
        build_feats32(acc, nf, feat32_written_serial(L), R.s1[0]);
        build_feats32(rc, nf, feat32_written_serial(L), R.s1[1]);
        in.s1[0] = R.s1[0].data(); in.s1[1] = R.s1[1].data(); in.nf1 = nf;
        in.s2 = E.s2p.data();
    }
    else
    {
        u32 nf = feat_count_read(L);
        build_feats(acc, nf, R.f1[0]);
        build_feats(rc, nf, R.f1[1]);
        in.f1[0] = R.f1[0].data(); in.f1[1] = R.f1[1].data(); in.nf1 = nf;
        in.f2 = E.f2p.data();
    }
    in.nf2 = E.nf2.data(); in.stop_ratio = E.stop_ratio;
    R.arena.resize(64 << 20);
    Arena ar = {R.arena.data(), R.arena.size(), 0, 0};
    int cap = 16 + (int)(L / 4);
    R.cords.assign(cap, 0);
    int nc = 0;
    PipeCounters cnt = {0, 0};
    u64 max_len = 0;
    for (int attempt = 0; attempt < 2; attempt++)
    {
        SeedTask t = make_task(acc, 0, 0, (u32)L, attempt ? 7 : 15);
        seed(E, acc, t, R.A);
        R.B.assign(R.A.size(), 0);
        const bool first_state = attempt == 0 && gdl_state;
        if (phase_c0(w, ar, R.hist, in, R.A.data(), R.B.data(), (int)R.A.size(), first_state ? 10 : 20, first_state ? 999 : 1, R.cords.data(), nc,
                     cap, cnt, max_len, true))
            return 1;
        if (!c0_attempt_too_short(max_len, L, in.win)) break;
    }
    c0_finish(w, L, R.cords.data(), nc, in.win);
    cords_out.assign(R.cords.begin(), R.cords.begin() + nc);
    return 0;
}

}  // namespace

extern "C" {

void * emu_create(int n_contigs, const uint8_t * const * dna5, const uint64_t * lens, int index_type, int feature_type,
                  int threads, int preset, int build_index)
{
    (void)index_type;
    Emu * E = new Emu();
    E->ft = feature_type == 1 ? 1 : 2;
    E->T = (unsigned)threads;
    E->stop_ratio = preset == 0 ? 0.7f : 0.0f;
    E->f2.resize(n_contigs);
    for (int i = 0; i < n_contigs; i++)
    {
        E->g.push_back(std::vector<u8>(dna5[i], dna5[i] + lens[i]));
        E->glen.push_back(lens[i]);
        SeqAcc acc = {E->g[i].data(), (i64)lens[i]};
        if (E->ft == 1) { E->s2.resize(n_contigs); build_feats32(acc, feat32_count(lens[i]), feat32_written_parallel(lens[i]), E->s2[i]); }
        else build_feats(acc, feat_count_genome(lens[i], E->T), E->f2[i]);
    }
    for (int i = 0; i < n_contigs; i++)
    {
        if (E->ft == 1) { E->s2p.push_back(E->s2[i].data()); E->nf2.push_back((u32)E->s2[i].size()); }
        else { E->f2p.push_back(E->f2[i].data()); E->nf2.push_back((u32)E->f2[i].size()); }
    }
    if (build_index) build_dindex(*E);
    return E;
}
void emu_destroy(void * h) { delete (Emu *)h; }
int64_t emu_dindex_dir(void * h, const int32_t ** p) { Emu * E = (Emu *)h; *p = E->dir.data(); return (int64_t)E->dir.size(); }
int64_t emu_dindex_hs(void * h, const uint64_t ** p) { Emu * E = (Emu *)h; *p = E->hs.data(); return (int64_t)E->hs.size(); }
int64_t emu_hindex_ysa(void *, const uint64_t ** p, uint64_t * e) { *p = 0; *e = 0; return 0; }
int64_t emu_hindex_dir_kv(void *, const uint64_t ** p, uint64_t * t) { *p = 0; *t = 0; return 0; }
int64_t emu_genome_features(void * h, int contig, const int32_t ** p)
{
    Emu * E = (Emu *)h;
    if (E->ft == 1)
    {
        E->outf.assign(E->s2[contig].begin(), E->s2[contig].end());
        *p = E->outf.data();
        return (int64_t)E->s2[contig].size();
    }
    *p = (const int32_t *)E->f2[contig].data();
    return (int64_t)E->f2[contig].size();
}
int64_t emu_read_features(void * h, const uint8_t * read, uint64_t len, int strand, const int32_t ** p)
{
    Emu * E = (Emu *)h;
    if (E->ft == 1)
    {
        std::vector<i16> f;
        if (strand) { RcAcc a = {read, (i64)len}; build_feats32(a, feat32_count(len), feat32_written_serial(len), f); }
        else { SeqAcc a = {read, (i64)len}; build_feats32(a, feat32_count(len), feat32_written_serial(len), f); }
        E->outf.assign(f.begin(), f.end());
        *p = E->outf.data();
        return (int64_t)f.size();
    }
    std::vector<F96> f;
    if (strand) { RcAcc a = {read, (i64)len}; build_feats(a, feat_count_read(len), f); }
    else { SeqAcc a = {read, (i64)len}; build_feats(a, feat_count_read(len), f); }
    E->outf.assign((const i32 *)f.data(), (const i32 *)f.data() + 3 * f.size());
    *p = E->outf.data();
    return (int64_t)f.size();
}
int64_t emu_read_stage(void * h, const uint8_t * read, uint64_t len, int stage, uint64_t str, uint64_t end, int toggle,
                       const uint64_t ** p)
{
    Emu * E = (Emu *)h;
    E->out.clear();
    SeqAcc acc = {read, (i64)len};
    if (stage == 1)
    {
        SeedTask t = make_task(acc, 0, (u32)str, (u32)end, toggle ? 7 : 15);
        seed(*E, acc, t, E->out);
    }
    else if (stage == 0) { if (run_read(*E, read, len, E->out, 0, 0)) E->out.assign(1, ~0ULL); }
    else if (stage == 4) { if (run_read(*E, read, len, E->out, 0, 1)) E->out.assign(1, ~0ULL); }
    else if (stage == 3)
    {
        std::vector<u64> c;
        if (run_read(*E, read, len, c, &E->out, 1)) E->out.assign(1, ~0ULL);
    }
    *p = E->out.data();
    return (int64_t)E->out.size();
}
int emu_map_batch(void * h, uint32_t n_reads, const uint8_t * bases, const uint64_t * read_off, int, uint64_t * cords,
                  uint64_t * cords_off, uint64_t cords_cap)
{
    Emu * E = (Emu *)h;
    u64 tot = 0;
    cords_off[0] = 0;
    for (uint32_t j = 0; j < n_reads; j++)
    {
        u64 len = read_off[j + 1] - read_off[j];
        std::vector<u64> c;
        if (len > (u64)kMinReadLen)
            if (run_read(*E, bases + read_off[j], len, c, 0, 0)) return -2;
        if (tot + c.size() > cords_cap) return -1;
        if (!c.empty()) std::memcpy(cords + tot, c.data(), 8 * c.size());
        tot += c.size();
        cords_off[j + 1] = tot;
    }
    return 0;
}

int emu_map_batch_c0(void * h, uint32_t n_reads, const uint8_t * bases, const uint64_t * read_off, int, int gdl_state, uint64_t * cords,
                     uint64_t * cords_off, uint64_t cords_cap)
{
    Emu * E = (Emu *)h;
    u64 tot = 0;
    cords_off[0] = 0;
    for (uint32_t j = 0; j < n_reads; j++)
    {
        u64 len = read_off[j + 1] - read_off[j];
        std::vector<u64> c;
        if (len > (u64)kMinReadLen)
            if (run_read_c0(*E, bases + read_off[j], len, c, gdl_state)) return -2;
        if (tot + c.size() > cords_cap) return -1;
        if (!c.empty()) std::memcpy(cords + tot, c.data(), 8 * c.size());
        tot += c.size();
        cords_off[j + 1] = tot;
    }
    return 0;
}

}  // extern "C"

// cords2BamLink through the product's walk (lnr_bamrec.h), count pass then fill pass, same output form as the checkers
extern "C" int64_t emu_cords2bam(void * h, uint64_t read_len, const uint64_t * cords, uint64_t n_cords, int window, uint64_t thd_large_X,
                                 int64_t thd_DI, int64_t thd_X, const int64_t ** recs, const uint64_t ** cigars, uint64_t * n_cigars)
{
    Emu * E = (Emu *)h;
    lnr::BamParms P = {(u32)window, thd_large_X, thd_DI, thd_X};
    lnr::BamCountSink cs;
    lnr::bam_walk(cords, (u32)n_cords, read_len, P, cs);
    std::vector<lnr::BamRec> r(cs.n_rec + 1);
    E->cigs.assign(cs.n_cig + 1, 0);
    lnr::BamFillSink fs;
    fs.recs = r.data(); fs.cig = E->cigs.data();
    lnr::bam_walk(cords, (u32)n_cords, read_len, P, fs);
    if (fs.n_rec != cs.n_rec || fs.n_cig != cs.n_cig) return -1;
    E->recs.clear();
    for (u32 k = 0; k < fs.n_rec; k++)
        for (int64_t v : {(int64_t)r[k].rid, (int64_t)r[k].begin_pos, (int64_t)r[k].flag, (int64_t)r[k].s1, (int64_t)r[k].s2, (int64_t)r[k].s3,
                          (int64_t)r[k].cigar_begin, (int64_t)r[k].cigar_end})
            E->recs.push_back(v);
    E->cigs.resize(fs.n_cig);
    *recs = E->recs.data(); *cigars = E->cigs.data(); *n_cigars = fs.n_cig;
    return (int64_t)fs.n_rec;
}

// gnu_sort must reproduce std::sort's permutation, ties included. Sorts `n` (key, payload) records with a
// key-only comparator by both and returns the number of positions that differ.
extern "C" int emu_gnu_sort_check(uint64_t * keys_payload, int n, int descending)
{
    std::vector<uint64_t> a(keys_payload, keys_payload + n), b(a);
    if (descending)
    {
        std::sort(a.begin(), a.end(), [](uint64_t & x, uint64_t & y) { return (x >> 32) > (y >> 32); });
        lnr::gnu_sort(b.data(), n, [](const uint64_t & x, const uint64_t & y) { return (x >> 32) > (y >> 32); });
    }
    else
    {
        std::sort(a.begin(), a.end(), [](uint64_t & x, uint64_t & y) { return (x >> 32) < (y >> 32); });
        lnr::gnu_sort(b.data(), n, [](const uint64_t & x, const uint64_t & y) { return (x >> 32) < (y >> 32); });
    }
    int bad = 0;
    for (int i = 0; i < n; i++) bad += a[i] != b[i];
    return bad;
}

// gnu_sort_w (the warp-cooperative statement: parallel Hoare partitions + stable radix pass) run with a one-lane warp
// must give std::sort's permutation too. Key = high 32 bits (30 significant), ascending in the transformed key.
struct EmuKeyHi { uint64_t operator()(uint64_t v) const { return v >> 32; } };
extern "C" int emu_gnu_sort_w_check(uint64_t * keys_payload, int n)
{
    std::vector<uint64_t> a(keys_payload, keys_payload + n), b(a), s0(n + 1), s1(n + 1);
    std::sort(a.begin(), a.end(), [](uint64_t & x, uint64_t & y) { return (x >> 32) < (y >> 32); });
    lnr::Warp w = {0, 1u};
    uint32_t hist[256];
    uint64_t * r = lnr::gnu_sort_w(w, hist, b.data(), s0.data(), s1.data(), n, 32, EmuKeyHi());
    int bad = 0;
    for (int i = 0; i < n; i++) bad += a[i] != r[i];
    return bad;
}

// std::sort itself (ascending by the high 32 bits), for the GPU self-test of gnu_sort_w
extern "C" void emu_std_sort_hi(uint64_t * a, int n)
{
    std::sort(a, a + n, [](uint64_t & x, uint64_t & y) { return (x >> 32) < (y >> 32); });
}
