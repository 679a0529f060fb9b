// Per-element building blocks shared by the CUDA kernels (lnr_kernels.cu) and the CPU-side logic check
// (tests/host_emu): index chunk geometry, one index sample, one 16-base feature cell, one seeding sample.
#pragma once
#include "lnr_defs.h"
#include "lnr_hash.h"

namespace lnr {

// ---- DIndex chunk geometry (createDIndex, index_util.cpp:1654-1670) ----------------------------------
// Contig of length len, T = threads_sem: chunk c rolls j over [tb[c]+span, tb[c+1]-span), tb[c] = len/T*c,
// tb[T] = len - span; a sample is taken at every 9th j starting at t_str + 8.
struct IdxChunk
{
    u64 base_off;   // offset of the contig in the concatenated genome buffer
    i64 len;        // contig length
    i64 t_str;      // first rolled position
    i64 n_samples;
    i64 kskip;      // hashInit's N-skip at t_str (0 for ACGT-only starts)
    i32 bias;       // strand-selector bias caused by kskip
    u32 contig;
    u64 sample0;    // global index of the chunk's first sample
};

LNR_HD void idx_chunk_range(i64 len, unsigned T, unsigned c, i64 & t_str, i64 & n_samples)
{
    i64 tb0 = len / (i64)T * (i64)c;
    i64 tb1 = (c + 1 == T) ? len - kSpanD : len / (i64)T * (i64)(c + 1);
    t_str = tb0 + kSpanD;
    i64 t_end = tb1 - kSpanD;
    i64 first = t_str + kIdxMinStep;
    n_samples = first < t_end ? (t_end - first + 8) / 9 : 0;
}

// one sample m of a chunk -> minimizer X and the 64-bit hs record (index_util.cpp:1763)
template <class BaseFn>
LNR_HD void idx_sample(BaseFn base, const IdxChunk & ch, i64 m, u32 & X, u64 & rec)
{
    i64 j = ch.t_str + kIdxMinStep + 9 * m;
    SeedVal sv;
    eval_sample<kSpanD>(base, j, j - ch.t_str + 1, ch.t_str + ch.kskip, ch.t_str + kSpanD - 1, ch.bias, sv);
    X = sv.X;
    rec = create_cord(ch.contig, (u64)j + kAnchorZero, sv.Y, sv.strand);
}

// ---- 2-mer / 48-base features (createFeatures2_48, pmpfinder.cpp:541-652) -----------------------------
// One 16-base cell: the 16 two-mers starting at p0 .. p0+15. Field id = 4a+b (TT and N add nothing);
// fields 0..9 live in `lo` (6 bits each), 10..14 in `hi`. An entry is the sum of three consecutive cells;
// no field exceeds 48, so nothing carries.
template <class BaseFn>
LNR_HD void feat_cell(BaseFn base, i64 p0, u64 & lo, u32 & hi)
{
    lo = 0; hi = 0;
    int a = base(p0);
#pragma unroll
    for (int i = 1; i <= 16; i++)
    {
        int b = base(p0 + i);
        if (a < 4 && b < 4)
        {
            int id = 4 * a + b;
            if (id < 10) lo += 1ULL << (6 * id);
            else if (id < 15) hi += 1u << (6 * (id - 10));
        }
        a = b;
    }
}
LNR_HD F96 feat_entry(u64 lo, u32 hi)
{
    F96 f;
    f.v[0] = (i32)(lo & ((1ULL << 30) - 1));
    f.v[1] = (i32)(lo >> 30);
    f.v[2] = (i32)hi;
    return f;
}
LNR_HD u32 feat_count_read(u64 L) { return L >= 50 ? (u32)((L - 50) / 16 + 1) : 0; }            // serial builder :556
LNR_HD u32 feat_count_genome(u64 len, unsigned T)                                               // parallel builder :589
{
    if (len < 48) return 0;
    u64 range = (len - 48) / 16 + 1;
    if (range < T) return feat_count_read(len);
    return (u32)(((len - 48) >> 4) + 1);
}

// ---- 1-mer / 32-base features (-f 1, createFeatures1_32 pmpfinder.cpp:354 serial / :393 parallel) ------------------
// one short per 16 bases: A + 32 C + 1024 G counts over the 32 bases from 16 i on. `written` = the entries the reference's
// builder writes; the entries behind them, and everything past the end of a string, read as 0 (canonical rule).
LNR_HD u32 feat32_count(u64 L) { return L >= 32 ? (u32)(((L - 32) >> 4) + 1) : 0; }
LNR_HD u32 feat32_written_serial(u64 L) { if (L < 32) return 0; u64 lim = L - 32; return (u32)(1 + (lim > 16 ? (lim - 16 + 15) / 16 : 0)); }
LNR_HD u32 feat32_written_parallel(u64 L) { return L >= 48 ? (u32)((L - 48) / 16) : 0; }
template <class BaseFn>
LNR_HD i16 feat32_entry(BaseFn base, i64 p0)
{
    int v = 0;
    for (int k = 0; k < 32; k++) { int b = base(p0 + k); v += b == 0 ? 1 : (b == 1 ? 32 : (b == 2 ? 1024 : 0)); }
    return (i16)v;
}

// ---- seeding task (getDIndexMatchAll, pmpfinder.cpp:1856): samples k = str + span + alpha*m - 1, m >= 1 ----
struct SeedTask
{
    u32 read;       // read index in the batch
    u32 str, end;   // [str, end) of the read
    u32 alpha;      // 15, or 7 on re-map
    u32 n_samples;
    i32 bias;       // 2 * (sum read[kskip..kskip+19] - sum read[k0..k0+19])
    u32 kskip;      // hashInit N-skip at the start of the read
    u32 pad;
    u64 sample0;    // global index of the task's first sample
};
LNR_HD u32 seed_task_samples(u32 str, u32 end, u32 alpha)
{
    i64 d = (i64)end - 2 * kSpanD - (i64)str;
    return d > 0 ? (u32)(d / alpha) : 0;
}
// getHIndexMatchAll (pmpfinder.cpp:1934): samples k = str + alpha*m - 1, m >= 1, while k < end - 17
LNR_HD u32 hseed_task_samples(u32 str, u32 end, u32 alpha)
{
    i64 d = (i64)end - 17 - (i64)str;
    return d > 0 ? (u32)(d / alpha) : 0;
}
template <class BaseFn>
LNR_HD void seed_sample(BaseFn base, const SeedTask & t, u32 m /* 1-based */, SeedVal & sv, u32 & k)
{
    i64 k0 = (i64)t.str + kSpanD;
    i64 kk = k0 + (i64)t.alpha * m - 1;
    eval_sample<kSpanD>(base, kk, kk - k0 + 1, (i64)t.kskip, k0 + kSpanD - 1, t.bias, sv);
    k = (u32)kk;
}

}  // namespace lnr
