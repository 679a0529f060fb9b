// State-free evaluation of the reference's rolling double-strand minimizer hash at one sample position.
//
// The reference rolls LShape over every base (shape_extend.cpp:86 hashInit, :173 hashNexth) and evaluates
// hashNextX (:341 = hashNextXX :245 + hashNextXY2 :282) at the sample positions. After >= span steps the
// forward / reverse-complement hashes are pure functions of the trailing `span` bases, and the strand
// selector `x` is the window sum plus a per-call constant. That lets every sample
// be evaluated independently by one GPU thread:
//
//   h  = sum W[w] * 4^(20-w)  mod 2^42        (an N, ord 4, carries into the previous base like the reference)
//   cr = sum ((3 - W[w]) & 3) << 2w           (an N contributes binary 11)
//   x  = 2 * sum s[p..p+20] - 63 + bias       bias = 2*(sum of hashInit's 20 bases - sum s[k0..k0+19])
//   strand = x > 0 ? 0 : 1 ; v2 = strand ? cr : h
//   X  = leftmost minimal 13-mer of v2 (26 bit), off = its offset 0..8
//   Y  = 4 flank bases after the minimizer on the chosen strand (8 bit), N / out-of-range -> 00
//
// For the first < span steps after hashInit the window is a splice of hashInit's bases and the newly fed
// bases (read seeding always, pmpfinder.cpp:1870-1874; index chunks only when hashInit skipped an N).
#pragma once
#include "lnr_defs.h"

namespace lnr {

struct SeedVal
{
    u32 X;       // minimizer, 2*weight bits
    u32 Y;       // 8-bit flank key
    u32 strand;  // 0 forward, 1 reverse complement
};

// BaseFn: (i64 pos) -> int ordinal 0..4; must return 0 for positions outside the sequence.
//   p        true position of the window start (j in createDIndex, k in getDIndexMatchAll)
//   n_steps  number of hashNexth calls since hashInit (>= 21: pure window)
//   init0    position of hashInit's first base (call start + N-skip)
//   feed0    position of the first base fed by hashNexth (k0 + span - 1)
// FULL_Y = false: Y is the 8-bit flank key of hashNextXY2 (index -i 1, all read seeding).
// FULL_Y = true : Y is hashNext's key (shape_extend.cpp:132, HIndex build): the bits of v2 outside the minimizer plus
//                 the minimizer offset code t << (2*span - 2*weight - 1), t = 64 - 2*span + 2*off.
template <int SPAN, bool FULL_Y, class BaseFn>
LNR_HD void eval_sample_t(BaseFn base, i64 p, i64 n_steps, i64 init0, i64 feed0, int bias, SeedVal & out)
{
    const int W = SPAN - 8;
    u64 h = 0, cr = 0;
    int sum = 0;
    if (n_steps >= SPAN)
    {
#pragma unroll
        for (int w = 0; w < SPAN; w++)
        {
            u64 c = (u64)base(p + w);
            h = (h << 2) + c;
            cr |= ((3 - c) & 3) << (2 * w);
            sum += (int)c;
        }
    }
    else
    {
        int n_init = SPAN - (int)n_steps;   // bases still coming from hashInit's window
        for (int w = 0; w < SPAN; w++)
        {
            i64 pos = w < n_init ? init0 + (n_steps - 1) + w : feed0 + (w - n_init);
            u64 c = (u64)base(pos);
            h = (h << 2) + c;
            cr |= ((3 - c) & 3) << (2 * w);
            sum += base(p + w);
        }
    }
    h &= (1ULL << (2 * SPAN)) - 1;
    int x = 2 * sum - 3 * SPAN + bias;
    u32 strand = x > 0 ? 0u : 1u;
    u64 v2 = strand ? cr : h;
    const u64 MX = (1ULL << (2 * W)) - 1;
    u64 X = ~0ULL;
    int off = 0;
#pragma unroll
    for (int o = 0; o <= 8; o++)
    {
        u64 v1 = (v2 >> (2 * (8 - o))) & MX;
        if (X > v1) { X = v1; off = o; }
    }
    u32 Y = 0;
    if (FULL_Y)
    {
        int below = 2 * (8 - off);
        u64 t = (u64)(64 - 2 * SPAN + 2 * off);
        Y = (u32)(((v2 >> (below + 2 * W)) << below) + (v2 & ((1ULL << below) - 1)) + (t << (2 * SPAN - 2 * W - 1)));
    }
    else if (!strand)
    {
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
            int c = base(p + W + off + q);
            Y = c > 3 ? (Y << 2) : (Y << 2) + (u32)c;
        }
    }
    else
    {
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
            int c = 3 - base(p + 7 - off - q);
            Y = c < 0 ? (Y << 2) : (Y << 2) + (u32)c;
        }
    }
    out.X = (u32)X;
    out.Y = Y;
    out.strand = strand;
}

template <int SPAN, class BaseFn>
LNR_HD void eval_sample(BaseFn base, i64 p, i64 n_steps, i64 init0, i64 feed0, int bias, SeedVal & out)
{
    eval_sample_t<SPAN, false>(base, p, n_steps, init0, feed0, bias, out);
}

// hashInit's N-skip (shape_extend.cpp:96-105): smallest k such that `span` consecutive non-N bases
// follow start+k. Returns k. (The scan is unbounded in the reference; `limit` guards the device code.)
template <int SPAN, class BaseFn>
LNR_HD i64 hash_init_skip(BaseFn base, i64 start, i64 limit)
{
    i64 k = 0, count = 0;
    while (count < SPAN && k + count < limit)
    {
        if (base(start + k + count) == 4) { k += count + 1; count = 0; }
        else count++;
    }
    return k;
}

// bias of the strand selector for a call that ran hashInit at init0 and started rolling at k0
template <int SPAN, class BaseFn>
LNR_HD int selector_bias(BaseFn base, i64 init0, i64 k0)
{
    int a = 0, b = 0;
    for (int i = 0; i < SPAN - 1; i++) { a += base(init0 + i); b += base(k0 + i); }
    return 2 * (a - b);
}

// Y-key match rule of getDIndexMatchAll (pmpfinder.cpp:1893): `val >> ctz(val) < 4`, ctz(0) accepted
LNR_HD bool ykey_match(u32 hs_y, u32 Y)
{
    u32 v = hs_y ^ Y;
    if (v == 0) return true;
#ifdef __CUDA_ARCH__
    return (v >> (__ffs((int)v) - 1)) < 4;
#else
    return (v >> __builtin_ctz(v)) < 4;
#endif
}

// DIndex::val2Anchor (index_util.cpp:1509)
LNR_HD u64 val2anchor(u64 hs, u64 k, u64 read_len, u32 strand)
{
    u64 hy = hs & kMaskY;
    if (cord_strand(hs) ^ strand)
    {
        u64 cy = read_len - 1 - k;
        return (hs - (cy << 20) + cy - hy) | kFlagStrand;
    }
    return (hs - (k << 20) + k - hy) & ~kFlagStrand;
}

}  // namespace lnr
