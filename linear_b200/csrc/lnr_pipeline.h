// Per-read approximate-map pipeline: everything apxMap (pmpfinder.cpp:2709) does after seeding.
//
// One warp owns one read. Data-parallel pieces (binning, radix sorts, the chaining DP's predecessor scan,
// traceback maxima, hit window filtering) are spread over the 32 lanes; the small data-dependent
// sequential pieces (run filter, traceback bookkeeping, block cutting, window extension, block chaining)
// run on lane 0 over the warp's private scratch arena. All comparator-tie behaviour of the reference's
// std::sort call sites is reproduced with lnr::gnu_sort (lnr_sortlib.h).
//
// The file is plain C++ when compiled without nvcc (single-lane warp, tests/host_emu) so the logic can be
// verified against the oracle on a CPU-only machine. Reference citations: file:line in the reference tree.
#pragma once
#include "lnr_defs.h"
#include "lnr_sortlib.h"
#include "lnr_warp.h"

namespace lnr {

// ----------------------------------------------------------------------------------------------------
// small types
// ----------------------------------------------------------------------------------------------------
struct ChainRec { i32 score, score2, len, p2anchor, root_ptr, f_leaf; };   // cluster_util.h:38-48
struct Blk { u32 first, second; };                                         // UPair of pointers [first, second)
struct YPair { u64 first, second; };                                       // UPair of cords

struct PipeIn
{
    const u8 * read;            // forward read bases (unused after seeding, kept for debugging)
    u32 L;                      // read length
    const F96 * f1[2];          // read features, strand 0 / 1           (createFeatures2_48 serial)
    u32 nf1;                    // entries per strand
    const F96 * const * f2;     // genome features per contig            (createFeatures2_48 parallel)
    const u32 * nf2;            // entries per contig
    // -f 1 (1-mer / 32-base features, one short per 16 bases): the same strings as shorts; ft selects the set
    const i16 * s1[2];
    const i16 * const * s2;
    int ft;                     // feature type: 2 = 2_48 (default), 1 = 1_32
    u32 win;                    // window size of the feature type: 96 / 192 (getFeatureWindowSize pmpfinder.cpp:244)
    float stop_ratio;           // ChainAnchorsHitsParms::thd_stop_chain_len_ratio (0.7 or 0, mapper.cpp:184)
};

// per-warp counters; t[] = clock64 cycles spent per pipeline stage on lane 0 (profiling aid, see kStageNames)
struct PipeCounters { u64 hits, windows; u64 t[16]; };
#ifdef __CUDA_ARCH__
#define LNR_CLOCK() clock64()
#else
#define LNR_CLOCK() 0LL
#endif
#define LNR_LAP(cnt, id, t0) do { long long t1_ = LNR_CLOCK(); (cnt).t[id] += (u64)(t1_ - (t0)); (t0) = t1_; } while (0)
// 0 binning, 1 sort ascending, 2 run filter, 3 sort by x (+ tie fallback), 4 chaining DP, 5 traceback, 6 hit blocks,
// 7 hit window filter, 8 path / window extension, 9 mid (clean, gaps), 10 finish (cord block chaining),
// 11 reads that needed the sequential tie-order sort, 12 reads

static const int kNumBins = (1 << 30) / 30000 + 2;   // binningFilter bins over a 30-bit x (pmpfinder.cpp:1984)

// ----------------------------------------------------------------------------------------------------
// window distance (pmpfinder.cpp:493-535)
// ----------------------------------------------------------------------------------------------------
LNR_HD int script_dist(i32 s1, i32 s2)   // __scriptDist63_31: packed 6-bit fields, bias 31, borrows kept
{
    const i32 mxu31 = (31 << 24) + (31 << 18) + (31 << 12) + (31 << 6) + 31;
    i32 d = (i32)((u32)s1 + (u32)mxu31 - (u32)s2);
    i32 a = ((d >> 24) & 63) - 31, b = ((d >> 18) & 63) - 31, c = ((d >> 12) & 63) - 31, e = ((d >> 6) & 63) - 31,
        f = (d & 63) - 31;
    return (a < 0 ? -a : a) + (b < 0 ? -b : b) + (c < 0 ? -c : c) + (e < 0 ? -e : e) + (f < 0 ? -f : f);
}
LNR_HD u32 window_dist48(const F96 * a, const F96 * b)   // _windowDist2_48: scripts at offsets {0,3}
{
    int s = 0;
#pragma unroll
    for (int i = 0; i < 6; i += 3)
        s += script_dist(a[i].v[0], b[i].v[0]) + script_dist(a[i].v[1], b[i].v[1]) + script_dist(a[i].v[2], b[i].v[2]);
    return (u32)s;
}
// -f 1: __scriptDist16_3 (pmpfinder.cpp:332): three 5-bit counters in a short, the top one by arithmetic shift
LNR_HD u32 script_dist16(i16 s1, i16 s2)
{
    int a = (s1 & 31) - (s2 & 31), b = ((s1 >> 5) & 31) - ((s2 >> 5) & 31), c = (s1 >> 10) - (s2 >> 10);
    return (u32)((a < 0 ? -a : a) + (b < 0 ? -b : b) + (c < 0 ? -c : c));
}
// _windowDist1_32 (:342): 6 scripts, every 2nd entry. Canonical rule for what the reference leaves undefined (it reads
// up to 10 entries past the end of a string, :693-703 / :924): an entry outside the string is 0
LNR_HD u32 window_dist32(const i16 * a, u32 na, u64 y, const i16 * b, u32 nb, u64 x)
{
    u32 d = 0;
#pragma unroll
    for (int i = 0; i < 12; i += 2)
        d += script_dist16(y + i < na ? a[y + i] : (i16)0, x + i < nb ? b[x + i] : (i16)0);
    return d;
}
// __windowDist (pmpfinder.cpp:655). The reference does not bounds-check here; for in-spec inputs the
// indices are in range, the guard only keeps the device from faulting on out-of-spec data.
LNR_HD u32 wdist(const PipeIn & in, u32 strand, u32 id, u64 y, u64 x, PipeCounters & cnt)
{
    cnt.windows++;
    if (y + 3 >= in.nf1 || x + 3 >= in.nf2[id]) return 1000;
    return window_dist48(in.f1[strand] + y, in.f2[id] + x);
}

// ----------------------------------------------------------------------------------------------------
// stable LSD radix sort of 64-bit keys by an extracted sub-key, 8 bits per pass, warp-cooperative.
// `src` is never written; passes ping-pong between t0 and t1; returns the buffer holding the result.
// Passes whose digit is identical for all keys are skipped. hist: 256 u32 of warp-private scratch.
// ----------------------------------------------------------------------------------------------------
struct KeyAsc { LNR_HD u64 operator()(u64 v) const { return v; } };
struct KeyXDesc { LNR_HD u64 operator()(u64 v) const { return (u64)0x3fffffffULL - anchor_x(v); } };
// One key functor for every sort of the pipeline, selected at run time: each functor type would instantiate its own
// copy of radix_sort / gnu_sort_w, and the section kernels are bound by instruction fetch as much as by anything else.
struct SortKey
{
    int mode;          // 0 anchor value ascending, 1 AnchorX descending, 2 y of the hit a cut points to, 3 Blk.second ascending
                       // (Blk viewed as u64: first low, second high), 4 x40 of a block's first record descending (elements
                       // are block indices), 5 / 6: the two sorts of the -c 0 path
    const u64 * recs;  // modes 2, 4
    const Blk * sep;   // mode 4
    LNR_HD u64 operator()(u64 v) const
    {
        switch (mode)
        {
        case 0: return v;
        case 1: return (u64)0x3fffffffULL - anchor_x(v);
        case 2: return cord_y(recs[v & ~(1ULL << 62)]);
        case 3: return v >> 32;
        case 5: return v & kMaskY;              // -c 0: Anchors::sortPos2 (base.cpp:303)
        case 6: return ~v;                      // -c 0: the run list, descending (getDHitList pmpfinder.cpp:2252)
        default: return ((1ULL << 40) - 1) - cord_x40(recs[sep[(u32)v].first]);
        }
    }
};
LNR_HD SortKey sort_key(int mode, const u64 * recs = (const u64 *)0, const Blk * sep = (const Blk *)0)
{
    SortKey k; k.mode = mode; k.recs = recs; k.sep = sep; return k;
}

template <class KeyFn>
LNR_PIPE u64 * radix_sort(const Warp & w, u32 * hist, u64 * src, u64 * t0, u64 * t1, int n, int key_bits, KeyFn key)
{
    // which bits vary at all
    u64 vor = 0, vand = ~0ULL;
    for (int i = w.lane; i < n; i += w.nl) { u64 k = key(src[i]); vor |= k; vand &= k; }
    vor = wor64(w, vor);
    vand = ~wor64(w, ~vand);
    u64 diff = vor ^ vand;
    u64 * in = src;
    u64 * out = t0;
    for (int shift = 0; shift < key_bits; shift += 8)
    {
        if (((diff >> shift) & 0xff) == 0) continue;
        for (int b = w.lane; b < 256; b += w.nl) hist[b] = 0;
        wsync(w);
        // four loads in flight per lane: the loop is bound by the round trip of in[i], not by the counting
        for (int c = 0; c < n; c += 4 * w.nl)
        {
            u64 v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) { int i = c + k * w.nl + w.lane; v[k] = i < n ? in[i] : 0; }
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                int i = c + k * w.nl + w.lane;
                if (i < n)
                {
                    u32 d = (u32)(key(v[k]) >> shift) & 0xff;
#ifdef __CUDA_ARCH__
                    atomicAdd(&hist[d], 1u);
#else
                    hist[d]++;
#endif
                }
            }
        }
        wsync(w);
        // exclusive scan of the 256 bins
        {
            int per = 256 / w.nl;
            int b0 = w.lane * per, s = 0;
            for (int b = 0; b < per; b++) s += (int)hist[b0 + b];
            int total;
            int base = wscan_excl(w, s, total);
            for (int b = 0; b < per; b++) { int c = (int)hist[b0 + b]; hist[b0 + b] = (u32)base; base += c; }
        }
        wsync(w);
        // stable scatter, 32 keys at a time in index order
        u64 v_next = w.lane < n ? in[w.lane] : 0;   // the next 32 keys are fetched while the current ones are ranked
        for (int c = 0; c < n; c += w.nl)
        {
            int i = c + w.lane;
            bool valid = i < n;
            u64 v = v_next;
            v_next = i + w.nl < n ? in[i + w.nl] : 0;
            u32 d = valid ? ((u32)(key(v) >> shift) & 0xff) : 0x100u + (u32)w.lane;
            u32 peers = wmatch(w, d);
            int rank = popc_below(w, peers);
            u32 base = valid ? hist[d] : 0;
            wsync(w);
            if (valid)
            {
                out[base + rank] = v;
                if (rank == popc32(peers) - 1) hist[d] = base + rank + 1;   // highest lane of the group
            }
            wsync(w);
        }
        in = out;
        out = (out == t0) ? t1 : t0;
    }
    wsync(w);
    return in;
}

// ----------------------------------------------------------------------------------------------------
// std::sort's permutation, warp-cooperative (lnr_sortlib.h is the one-lane statement of the same algorithm).
//
// less(x, y) = key(x) < key(y). libstdc++'s introsort is (1) a tree of Hoare partitions down to ranges of <= 16
// elements and (2) one final insertion sort. Both parts have a data-parallel equivalent that yields the same array:
//
// (1) __unguarded_partition(lo, hi, pivot) stops its left scan at elements >= pivot and its right scan at elements
//     <= pivot and swaps the k-th left stop with the k-th right stop until the scans cross. Every stop before the
//     crossing is an element of the ORIGINAL range (swapped elements are always stepped over), so with l_1 < l_2 < ..
//     the positions of the elements >= pivot and r_1 > r_2 > .. those of the elements <= pivot, the swaps are exactly
//     (l_k, r_k) for k = 1..K, K = #{k : l_k < r_k}, and the returned cut is l_1 when K = 0, else min(l_{K+1}, r_K)
//     (the left scan cannot pass r_K, which now holds an element >= pivot).
// (2) insertion sort moves an element left only past strictly greater ones: the result is the STABLE sort of the
//     array the partitions left behind, which the LSD radix sort above computes.
// The depth-limit heap sort (rare) stays on one lane.
// ----------------------------------------------------------------------------------------------------
template <class KeyFn>
LNR_PIPE int gs_partition_w(const Warp & w, u64 * a, int lo, int hi, int pivot, int * Lidx, int * Ridx, KeyFn key)
{
    const u64 pv = key(a[pivot]);
    const int m = hi - lo;
    int nL = 0, nR = 0;
    for (int c = 0; c < m; c += w.nl)
    {
        int i = lo + c + w.lane, j = hi - 1 - c - w.lane;
        bool vi = i < hi, vj = j >= lo;
        u64 ki = vi ? key(a[i]) : 0, kj = vj ? key(a[j]) : 0;
        bool isL = vi && !(ki < pv);
        bool isR = vj && !(pv < kj);
        u32 bl = wballot(w, isL), br = wballot(w, isR);
        if (isL) Lidx[nL + popc_below(w, bl)] = i;
        if (isR) Ridx[nR + popc_below(w, br)] = j;
        nL += popc32(bl);
        nR += popc32(br);
    }
    wsync(w);
    const int mn = nL < nR ? nL : nR;
    int K = 0;
    for (int c = 0; c < mn; c += w.nl)
    {
        int k = c + w.lane;
        bool p = k < mn && Lidx[k] < Ridx[k];
        int cnt = popc32(wballot(w, p));
        K += cnt;
        if (cnt < w.nl) break;
    }
    int cut;
    if (K == 0) cut = Lidx[0];
    else
    {
        int rK = Ridx[K - 1];
        cut = (K < nL && Lidx[K] < rK) ? Lidx[K] : rK;
    }
    for (int k = w.lane; k < K; k += w.nl)
    {
        int x = Lidx[k], y = Ridx[k];
        u64 t = a[x]; a[x] = a[y]; a[y] = t;
    }
    wsync(w);
    return cut;
}

// sorts a[0..n) like std::sort(a, a + n, less); s0 / s1: scratch of n u64 each. Returns the buffer holding the result
// (a, s0 or s1).
template <class KeyFn>
LNR_PIPE u64 * gnu_sort_w(const Warp & w, u32 * hist, u64 * a, u64 * s0, u64 * s1, int n, int key_bits, KeyFn key)
{
    if (n <= 1) return a;
    auto less = [key](const u64 & x, const u64 & y) { return key(x) < key(y); };
    if (n <= 16)
    {
        if (w.lane == 0) gs_insertion_sort(a, 0, n, less);
        wsync(w);
        return a;
    }
    int lg = 0;
    for (unsigned v = (unsigned)n; v > 1; v >>= 1) lg++;
    int * Lidx = (int *)s0;
    int * Ridx = (int *)s1;
    int st_first[72], st_last[72], st_depth[72];
    int sp = 1;
    st_first[0] = 0; st_last[0] = n; st_depth[0] = lg * 2;
    while (sp > 0)
    {
        --sp;
        int first = st_first[sp], last = st_last[sp], depth = st_depth[sp];
        while (last - first > 16)
        {
            if (depth == 0)
            {
                if (w.lane == 0) gs_heap_sort(a, first, last, less);
                wsync(w);
                break;
            }
            --depth;
            int mid = first + (last - first) / 2;
            if (w.lane == 0) gs_move_median_to_first(a, first, first + 1, mid, last - 1, less);
            wsync(w);
            int cut = gs_partition_w(w, a, first + 1, last, first, Lidx, Ridx, key);
            st_first[sp] = cut; st_last[sp] = last; st_depth[sp] = depth; sp++;
            last = cut;
        }
    }
    return radix_sort(w, hist, a, s0, s1, n, key_bits, key);
}

// ----------------------------------------------------------------------------------------------------
// anchor filters (pmpfinder.cpp:1979-2183)
// The same filter on all lanes. The recurrence is sequential only through its breaks: inside a run the median the next
// anchor is tested against is a[(block_str + i - 1) >> 1], known without looking at the decisions before it. So every lane
// tests one of the next anchors ON THE ASSUMPTION that all anchors between the round's first and its own continue the run;
// the lanes below the first break (or the last anchor) were right and are folded in at once, the break itself is applied
// as the reference does, and the next round starts behind it. A clean read advances 32 anchors per round, a noisy stretch
// one break per round. One lane (host build) degenerates to the sequential loop of filter_anchor_runs below.
LNR_PIPE int filter_anchor_runs_w(const Warp & w, const u64 * a, int n, Blk * ranges)
{
    int nr = 0;
    if (n < 2) return 0;
    u64 ak2 = a[1];
    u64 block_str = 1, count = 0, min_y = ~0ULL, max_y = 0;
    int i0 = 1;
    while (i0 < n)
    {
        const int i = i0 + w.lane;
        const bool valid = i < n;
        u64 ai = 0, ref = ak2;
        if (valid)
        {
            ai = a[i];
            if (w.lane > 0) ref = a[(block_str + (u64)(i - 1)) >> 1];
        }
        const u64 y = cord_y(ai);
        const u64 dy2 = (u64)iabs64((i64)(y - cord_y(ref)));
        const bool cont = valid && cord_x40(ai - ref) < (dy2 >> 2);
        const u32 ev = wballot(w, valid && (!cont || i == n - 1));
        const int f = ev ? ffs32(ev) : w.nl;       // lanes below f continue the run (the last anchor is always an event)
        if (f > 0)
        {
            const u32 mn = ~wmax_u32(w, w.lane < f ? ~(u32)y : 0u), mx = wmax_u32(w, w.lane < f ? (u32)y : 0u);
            if (min_y > mn) min_y = mn;
            if (max_y < mx) max_y = mx;
            count += (u64)f;
            ak2 = a[(block_str + (u64)(i0 + f - 1)) >> 1];
        }
        if (ev)
        {
            const int ix = i0 + f;
            const u64 yf = (u64)wbcast(w, (u32)y, f);
            const u64 af = wbcast64(w, ai, f);
            if ((wballot(w, cont) >> f) & 1u)      // the last anchor, continuing the run
            {
                if (min_y > yf) min_y = yf;
                if (max_y < yf) max_y = yf;
                ++count;
            }
            const u64 thd = umax64((max_y - min_y) >> 10, 2);
            if (count > thd)
            {
                if (w.lane == 0) { ranges[nr].first = (u32)block_str; ranges[nr].second = (u32)ix; }
                nr++;
            }
            block_str = (u64)ix;
            ak2 = af;
            min_y = yf;
            max_y = yf;
            count = 1;
            i0 = ix + 1;
        }
        else i0 += w.nl;
    }
    wsync(w);
    return nr;
}

// ----------------------------------------------------------------------------------------------------
LNR_HD u32 anchor_bin(u64 a) { return (u32)(cord_x(a) / 30000); }

// Per-warp scratch in shared memory: the radix histogram (256) followed by the bin sketch of binning_filter.
static const int kSketchBits = 10;
static const int kSketch = 1 << kSketchBits;
static const int kWarpSmemWords = 256 + kSketch;
LNR_HD u32 sketch_slot(u32 bin) { return (bin * 2654435761u) >> (32 - kSketchBits); }

// binningFilter (:1979). A[0..n) -> survivors in B (or A unchanged when nothing survives). Returns the
// buffer that holds the result and its length through n_out. `bins` is the warp's zeroed histogram (one counter per
// 30-kb bin of the 30-bit x axis, 140 KB) and is returned zeroed; `sketch` (kSketch zeroed counters in shared memory)
// likewise.
//
// The histogram of a warp is far larger than its share of L2, so every access to it is a DRAM round trip, while nearly
// all anchors of a read are chance hits alone in their bin. The sketch counts hashed bins in shared memory first: a
// sketch counter is >= the true count of each bin that maps to it, so an anchor whose counter is <= 10 can be dropped
// without looking at the histogram at all, and only the few candidates take the exact, global path.
// Two passes over A (sketch; exact counts of the candidates, which are copied to B in order), then two over the short
// candidate list only: mark the candidates whose bin holds > 10 (bit 63 of the copy: raw anchors never carry it, their
// contig id ends at bit 59 and only the strand bit 61 lies above), then clear the candidates' bins and compact the marked
// ones in place. (Four passes over A took 48 % of k_hits_sort.) When most anchors are candidates the list would only add
// copies, and the passes walk A itself.
LNR_PIPE u64 * binning_filter(const Warp & w, u32 * bins, u32 * sketch, u64 * A, u64 * B, int n, int & n_out)
{
    const int U = 4, step = U * w.nl;
    for (int c = 0; c < n; c += step)
    {
        u64 v[U];
#pragma unroll
        for (int k = 0; k < U; k++) { int i = c + k * w.nl + w.lane; v[k] = i < n ? A[i] : 0; }
#pragma unroll
        for (int k = 0; k < U; k++)
        {
            int i = c + k * w.nl + w.lane;
            if (i < n)
            {
#ifdef __CUDA_ARCH__
                atomicAdd(&sketch[sketch_slot(anchor_bin(v[k]))], 1u);
#else
                sketch[sketch_slot(anchor_bin(v[k]))]++;
#endif
            }
        }
    }
    wsync(w);
    // the number of candidates is known from the sketch alone: the sum of the counters above 10
    int n_cand = 0;
    for (int j = w.lane; j < kSketch; j += w.nl) { u32 v = sketch[j]; if (v > 10) n_cand += (int)v; }
    n_cand = wsum(w, n_cand);
    if (2 * n_cand > n)
    {
        // dense case (repeat-rich reads, HIndex seeding: most anchors are candidates): a candidate list would only add
        // copies -- exact counts, compaction and clearing walk A itself
        for (int c = 0; c < n; c += step)
        {
            u64 v[U];
#pragma unroll
            for (int k = 0; k < U; k++) { int i = c + k * w.nl + w.lane; v[k] = i < n ? A[i] : 0; }
#pragma unroll
            for (int k = 0; k < U; k++)
            {
                int i = c + k * w.nl + w.lane;
                if (i < n)
                {
                    u32 bin = anchor_bin(v[k]);
                    if (sketch[sketch_slot(bin)] > 10)
                    {
#ifdef __CUDA_ARCH__
                        atomicAdd(&bins[bin], 1u);
#else
                        bins[bin]++;
#endif
                    }
                }
            }
        }
        wsync(w);
        int id = 0;
        for (int c = 0; c < n; c += step)
        {
            u64 v[U]; u32 cnt[U];
#pragma unroll
            for (int k = 0; k < U; k++) { int i = c + k * w.nl + w.lane; v[k] = i < n ? A[i] : 0; }
#pragma unroll
            for (int k = 0; k < U; k++)
            {
                int i = c + k * w.nl + w.lane;
                cnt[k] = 0;
                if (i < n)
                {
                    u32 bin = anchor_bin(v[k]);
                    if (sketch[sketch_slot(bin)] > 10) cnt[k] = bins[bin];
                }
            }
#pragma unroll
            for (int k = 0; k < U; k++)
            {
                bool keep = cnt[k] > 10;
                u32 bal = wballot(w, keep);
                if (keep) B[id + popc_below(w, bal)] = v[k];
                id += popc32(bal);
            }
        }
        wsync(w);
        for (int c = 0; c < n; c += step)
        {
            u64 v[U];
#pragma unroll
            for (int k = 0; k < U; k++) { int i = c + k * w.nl + w.lane; v[k] = i < n ? A[i] : 0; }
#pragma unroll
            for (int k = 0; k < U; k++)
            {
                int i = c + k * w.nl + w.lane;
                if (i < n)
                {
                    u32 bin = anchor_bin(v[k]);
                    if (sketch[sketch_slot(bin)] > 10) bins[bin] = 0;
                }
            }
        }
        wsync(w);
        for (int j = w.lane; j < kSketch; j += w.nl) sketch[j] = 0;
        wsync(w);
        if (id != 0) { n_out = id; return B; }
        n_out = n;
        return A;
    }
    int nc = 0;                                  // candidates, copied to B[0..nc) in the order of A
    for (int c = 0; c < n; c += step)
    {
        u64 v[U];
#pragma unroll
        for (int k = 0; k < U; k++) { int i = c + k * w.nl + w.lane; v[k] = i < n ? A[i] : 0; }
#pragma unroll
        for (int k = 0; k < U; k++)
        {
            int i = c + k * w.nl + w.lane;
            bool cand = false;
            if (i < n)
            {
                u32 bin = anchor_bin(v[k]);
                cand = sketch[sketch_slot(bin)] > 10;
                if (cand)
                {
#ifdef __CUDA_ARCH__
                    atomicAdd(&bins[bin], 1u);
#else
                    bins[bin]++;
#endif
                }
            }
            u32 bal = wballot(w, cand);
            if (cand) B[nc + popc_below(w, bal)] = v[k];
            nc += popc32(bal);
        }
    }
    wsync(w);
    for (int c = 0; c < nc; c += w.nl)             // mark
    {
        int i = c + w.lane;
        if (i < nc)
        {
            u64 v = B[i];
            if (bins[anchor_bin(v)] > 10) B[i] = v | kFlagMain;
        }
    }
    wsync(w);
    int ii = 0;
    for (int c = 0; c < nc; c += w.nl)             // clear the bins, keep the marked candidates (ii <= c: in place)
    {
        int i = c + w.lane;
        u64 v = i < nc ? B[i] : 0;
        if (i < nc) bins[anchor_bin(v)] = 0;
        wsync(w);
        bool keep = (v & kFlagMain) != 0;
        u32 bal = wballot(w, keep);
        if (keep) B[ii + popc_below(w, bal)] = v & ~kFlagMain;
        ii += popc32(bal);
        wsync(w);
    }
    for (int j = w.lane; j < kSketch; j += w.nl) sketch[j] = 0;
    wsync(w);
    if (ii != 0) { n_out = ii; return B; }
    n_out = n;
    return A;
}

// filterAnchorsList (:2019) on ascending-sorted anchors a[0..n), a[0] == 0. Writes accepted ranges and
// returns their number. Sequential (running median).
LNR_HD int filter_anchor_runs(const u64 * a, int n, Blk * ranges)
{
    int nr = 0;
    u64 ak2 = a[1];
    u64 block_str = 1, count = 0, min_y = ~0ULL, max_y = 0;
    for (int i = 1; i < n; i++)
    {
        u64 y = cord_y(a[i]);
        u64 dy2 = (u64)iabs64((i64)(y - cord_y(ak2)));
        bool cont = cord_x40(a[i] - ak2) < (dy2 >> 2);
        if (cont)
        {
            if (min_y > y) min_y = y;
            if (max_y < y) max_y = y;
            ak2 = a[(block_str + (u64)i) >> 1];
            ++count;
        }
        if (!cont || i == n - 1)
        {
            u64 thd = umax64((max_y - min_y) >> 10, 2);
            if (count > thd) { ranges[nr].first = (u32)block_str; ranges[nr].second = (u32)i; nr++; }
            block_str = (u64)i;
            ak2 = a[i];
            min_y = y;
            max_y = y;
            count = 1;
        }
    }
    return nr;
}

// ----------------------------------------------------------------------------------------------------
// chain scores (cluster_util.cpp:337-443, :586-860)
// ----------------------------------------------------------------------------------------------------
// The reference computes these in int64. y is a 20-bit and x a 30-bit coordinate, so dy, dx and da = |dx - dy| fit an
// int32 and only 100 * da can leave 32 bits (da >= 2^25): the 64-bit division (a long subroutine on the GPU, and the
// chaining DP evaluates a score per candidate) is kept for that rare case only.
LNR_HD u32 chain_derr(i32 da, i32 dy, i32 dx, i32 floor_)
{
    i32 ady = dy < 0 ? -dy : dy, adx = dx < 0 ? -dx : dx;
    i32 den = ady > adx ? ady : adx;
    if (den < floor_) den = floor_;
    if (da < (1 << 25)) return (100u * (u32)da) / (u32)den;
    return (u32)((100ULL * (u64)da) / (u64)den);
}
// (x, y) = (anchor_x, cord_y) of the two anchors, decoded once by the caller
LNR_HD int score_anchor_xy(i32 x1, i32 y1, i32 x2, i32 y2)   // getApxChainScore :387
{
    i32 dy = y1 - y2;
    if (dy < 10) return -10000;
    i32 dx = x1 - x2;
    i32 da = dx - dy;
    if (da < 0) da = -da;
    i32 derr = (i32)chain_derr(da, dy, dx, 50);
    int sderr;
    if (derr < 5) sderr = 4 * derr;
    else if (derr < 10) sderr = 6 * derr - 10;
    else if (derr < 100) sderr = derr * derr - 5 * derr;
    else return -1000;
    int sdy;
    dy /= 15;
    if (dy < 150) sdy = dy / 5;
    else if (dy < 100) sdy = dy - 30;
    else if (dy < 10000) sdy = dy * dy / 200 + 20;
    else sdy = 10000;
    return da < 10 ? 100 - sdy : 100 - sdy - sderr;
}
LNR_HD int score_anchor0_xy(i32 x1, i32 y1, i32 x2, i32 y2)   // getApxChainScore0 :337
{
    i32 dy = y1 - y2;
    if (dy < 5) return -10000;
    i32 dx = x1 - x2;
    i32 da = dx - dy;
    if (da < 0) da = -da;
    if (chain_derr(da, dy, dx, 50) >= 100) return -1000;
    int sdy = dy, sderr = da;
    return da < 30 ? 100 - sdy : 100 - sdy - sderr;
}
LNR_HD int score_anchor(u64 a1, u64 a2) { return score_anchor_xy((i32)anchor_x(a1), (i32)cord_y(a1), (i32)anchor_x(a2), (i32)cord_y(a2)); }
LNR_HD int score_anchor0(u64 a1, u64 a2) { return score_anchor0_xy((i32)anchor_x(a1), (i32)cord_y(a1), (i32)anchor_x(a2), (i32)cord_y(a2)); }
LNR_HD int score_blocks_hits(u64 c11, u64 c22)   // getApxChainScore2 :586
{
    i64 dy = (i64)(cord_y(c11) - cord_y(c22));
    i64 dx = (i64)(cord_x(c11) - cord_x(c22));
    if (dx < 0 || dy < 0 || cord_strand(c11 ^ c22) || dx > 20000 || dy > 20000) return (int)0x80000000;
    i64 da = iabs64(dx - dy);
    i64 derr = (100 * da) / imax64(imax64(iabs64(dy), 100), iabs64(dx));
    if (da > 100 || derr > 50)
    {
        if (dx < dy) return (int)(100 - 30 - dy / 1000 - dx / 100);
        return (int)(100 - 30 - dy / 100 - dx / 1000);
    }
    return (int)(100 - dy / 95);
}
LNR_HD int score_blocks_cords(u64 c11, u64 c12, u64 c21, u64 c22, u64 L, int strand)   // getApxChainScore3 :811
{
    i64 dx, dy;
    // getChainBlockDxDy :774
    if (cord_strand(c11) != (u64)strand)
    {
        if (cord_strand(c22) != (u64)strand) { dy = (i64)(cord_y(c21) - cord_y(c12)); dx = (i64)(cord_x(c21) - cord_x(c12)); }
        else { dy = (i64)(L - cord_y(c12) - 1 - cord_y(c22)); dx = (i64)(cord_x(c11) - cord_x(c22)); }
    }
    else
    {
        if (cord_strand(c22) != (u64)strand) { dy = (i64)(cord_y(c11) - L + 1 + cord_y(c21)); dx = (i64)(cord_x(c11) - cord_x(c22)); }
        else { dy = (i64)(cord_y(c11) - cord_y(c22)); dx = (i64)(cord_x(c11) - cord_x(c22)); }
    }
    int f_type = (int)cord_strand(c11 ^ c22);
    i64 min_dy = -80, min_dx = -(i64)L;
    i64 max_dy = (i64)((float)L * 1.0f);
    i64 max_dx = 15000, dup_trigger = -50;
    i64 dx_ = iabs64(dx), dy_ = iabs64(dy), da = dx - dy;
    int score = 0;
    if (dy < min_dy || dy > max_dy || dx < min_dx || dx_ > max_dx) score = (int)0x80000000;
    else
    {
        i64 sdy = dy_ > 2000 ? imin64(dy_ / 25 - 50, 70) : dy_ / 40;
        i64 sdx = dx_ > 2000 ? imin64(dx_ / 25 - 50, 70) : dx_ / 40;
        if (f_type == 1) { if (dx > min_dx) score = (int)(75 - sdy); }
        else if (da < -imax64(dx_ / 4, 50))
        {
            if (dx > dup_trigger) score = (int)(80 - sdx);
            else score = (int)(80 - sdy);
        }
        else if (da > imax64(dy / 4, 50)) score = (int)(80 - sdy);
        else score = (int)(100 - sdy);
    }
    return score;
}

// ----------------------------------------------------------------------------------------------------
// getBestChains (cluster_util.cpp:53): for each i the best predecessor among the previous 20 or any with
// x_j - x_i < 300. The reference scans j downward and overwrites on `>=`, so among equal best sums the
// smallest j wins; the scan stops at the first j that fails the range test.
//
// The recurrence over i is sequential. Lane l keeps predecessor j = i-1-l (anchor, chain score, chain
// length, root) in registers and the window slides by one shuffle per step, so the inner loop has no
// dependent global loads; predecessors older than the register window (dense repeats only) come from
// global memory. The winner is found with one 32-bit warp max over (sum << 5 | lane).
// ----------------------------------------------------------------------------------------------------
LNR_HD int score_pair(u64 aj, u64 ai, int score_type) { return score_type == 0 ? score_anchor(aj, ai) : score_anchor0(aj, ai); }

LNR_PIPE void best_chains(const Warp & w, const u64 * a, ChainRec * ch, int n, int score_type)
{
    const int depth = 20;
    const u64 dx_depth = 300;
    i32 xj = 0, yj = 0;               // anchor_x / cord_y of predecessor i-1-lane (decoded once, when it enters the window)
    i32 sj = 0, lenj = 0, rootj = 0;  // its chain score / length / root
    u64 a_next = n > 0 ? a[0] : 0;
    for (int i = 0; i < n; i++)
    {
        const u64 ai = a_next;
        if (i + 1 < n) a_next = a[i + 1];
        const u64 xi = anchor_x(ai);
        const i32 xi32 = (i32)xi, yi32 = (i32)cord_y(ai);
        const int j = i - 1 - w.lane;
        bool ok = j >= 0 && (w.lane < depth || (u64)(u32)xj - xi < dx_depth);
        u32 okmask = wballot(w, ok);
        u32 full = w.nl == 32 ? 0xffffffffu : ((1u << w.nl) - 1);
        int first_fail = okmask == full ? w.nl : ffs32(~okmask & full);
        u32 key = 0;                  // (sum << 5) | lane ; 0 = none
        if (ok && w.lane < first_fail)
        {
            int s = score_type == 0 ? score_anchor_xy(xj, yj, xi32, yi32) : score_anchor0_xy(xj, yj, xi32, yi32);
            if (s > 0) key = ((u32)(s + sj) << 5) | (u32)w.lane;
        }
        key = wmax_u32(w, key);
        int new_max = key ? (int)(key >> 5) : -1;
        int max_j = key ? i - 1 - (int)(key & 31) : i;
        i32 p_len = 0, p_root = 0;
        if (key) { p_len = wbcast(w, lenj, (int)(key & 31)); p_root = wbcast(w, rootj, (int)(key & 31)); }
        // predecessors beyond the register window: only when every window lane passed the range test
        if (first_fail == w.nl && i - 1 - w.nl >= 0)
        {
            i64 best = key ? (((i64)new_max << 32) | (i64)(0x7fffffff - max_j)) : -1;
            bool from_mem = false;
            for (int jb = i - 1 - w.nl; jb >= 0; jb -= w.nl)
            {
                int jj = jb - w.lane;
                bool ok2 = jj >= 0;
                u64 a2 = ok2 ? a[jj] : 0;
                ok2 = ok2 && (jj >= i - depth || anchor_x(a2) - xi < dx_depth);
                u32 m2 = wballot(w, ok2);
                int ff = m2 == full ? w.nl : ffs32(~m2 & full);
                if (ok2 && w.lane < ff)
                {
                    int s = score_pair(a2, ai, score_type);
                    if (s > 0)
                    {
                        i64 k2 = (((i64)s + (i64)ch[jj].score) << 32) | (i64)(0x7fffffff - jj);
                        if (k2 > best) best = k2;
                    }
                }
                if (ff != w.nl) break;
            }
            i64 bmax = wmax_i64(w, best);
            if (bmax >= 0)
            {
                int mj = 0x7fffffff - (int)(bmax & 0x7fffffff);
                from_mem = mj != max_j || !key;
                new_max = (int)(bmax >> 32);
                max_j = mj;
            }
            if (from_mem) { p_len = ch[max_j].len; p_root = ch[max_j].root_ptr; }
        }
        i32 c_score, c_len, c_root;
        if (new_max > 0) { c_score = new_max; c_len = p_len + 1; c_root = p_root; }
        else { c_score = 0; c_len = 1; c_root = i; max_j = -1; }
        if (w.lane == 0)
        {
            ChainRec r;
            r.p2anchor = max_j; r.score = c_score; r.score2 = c_score; r.len = c_len; r.root_ptr = c_root; r.f_leaf = 1;
            ch[i] = r;
            if (max_j >= 0) ch[max_j].f_leaf = 0;
        }
        // slide the window: lane l takes lane l-1, lane 0 takes element i
        xj = (i32)wshift_up32(w, (u32)xj); yj = (i32)wshift_up32(w, (u32)yj);
        sj = (i32)wshift_up32(w, (u32)sj); lenj = (i32)wshift_up32(w, (u32)lenj); rootj = (i32)wshift_up32(w, (u32)rootj);
        if (w.lane == 0) { xj = xi32; yj = yi32; sj = c_score; lenj = c_len; rootj = c_root; }
    }
    wsync(w);
}

// ----------------------------------------------------------------------------------------------------
// traceback (cluster_util.cpp:122-332). Chains are written back to back into out_el / out_score;
// chain_off[c] .. chain_off[c+1] delimit chain c. Sequential. Returns the number of chains.
// ----------------------------------------------------------------------------------------------------
template <class E>
LNR_HD_COLD int traceback0(const E * el, ChainRec * rec, int n, E * out_el, i32 * out_score, int * chain_off, int max_chains,
                      int min_len, int abort_score, int bestn, float stop_ratio)
{   // traceBackChains0 :122
    const int delete_score = -1000;
    int n_chains = 0, pos = 0;
    chain_off[0] = 0;
    int search_times = bestn < 50 ? bestn : 50;
    for (int it = 0; it < search_times; it++)
    {
        bool f_done = true;
        int max_2nd = -1, max_score = -1, max_str = -1, max_len = 0;
        for (int j = 0; j < n; j++)
            if (rec[j].score > max_score)
            {
                max_2nd = max_score; max_str = j; max_score = rec[j].score; max_len = rec[j].len; f_done = false;
            }
        if (n_chains > 0)
            if ((float)max_len > (float)(u64)(chain_off[1] - chain_off[0]) * stop_ratio) f_done = false;
        if (f_done || max_score == 0) break;
        if (max_len > min_len && max_score / (max_len - 1) > abort_score)
        {
            int cur = pos;   // tentative chain [pos, cur)
            for (int j = max_str; j != -1; j = rec[j].p2anchor)
            {
                if (rec[j].score != delete_score)
                {
                    out_el[cur] = el[j];
                    out_score[cur] = rec[j].score2;
                    cur++;
                    rec[j].score = delete_score;
                }
                else
                {
                    int infix = rec[j].score2;
                    if (max_score - infix < max_2nd)
                    {
                        for (int k = max_str; k != j; k = rec[k].p2anchor) rec[k].score = rec[k].score2 - infix;
                        cur = pos;
                    }
                    break;
                }
            }
            if (cur != pos && n_chains < max_chains)
            {
                pos = cur;
                n_chains++;
                chain_off[n_chains] = pos;
            }
        }
        if (max_str != -1) rec[max_str].score = delete_score;
    }
    return n_chains;
}

template <class E>
LNR_HD_COLD int traceback1(const E * el, ChainRec * rec, int n, E * out_el, i32 * out_score, int * chain_off, int max_chains,
                      int min_len, int abort_score, int bestn, float stop_ratio)
{   // traceBackChains1 :214 -- called only when there are <= 50 trees
    int root[64], lscore[64], llen[64], lidx[64];
    int nt = 0;
    for (int j = 0; j < n; j++)
        if (rec[j].f_leaf)
        {
            int f_new = 1;
            for (int k = 0; k < nt; k++)
                if (root[k] == rec[j].root_ptr)
                {
                    if (rec[j].score > lscore[k]) { lscore[k] = rec[j].score; llen[k] = rec[j].len; lidx[k] = j; }
                    f_new = 0;
                }
            if (f_new && nt < 64) { root[nt] = rec[j].root_ptr; lscore[nt] = rec[j].score; llen[nt] = rec[j].len; lidx[nt] = j; nt++; }
        }
    struct Rank { int first, second; };
    Rank ranks[64];
    for (int i = 0; i < nt; i++) { ranks[i].first = i; ranks[i].second = lscore[i]; }
    gnu_sort(ranks, nt, [](const Rank & a, const Rank & b) { return a.second > b.second; });
    int n_chains = 0, pos = 0, f_stop = 0;
    int stale = 0;   // the reference keeps appending to an uncleared `chain` once f_stop is set (dead data)
    chain_off[0] = 0;
    int lim = bestn < nt ? bestn : nt;
    for (int i = 0; i < lim; i++)
    {
        int t = ranks[i].first;
        int max_score = lscore[t], max_len = llen[t], max_str = lidx[t];
        int mean = max_len > 1 ? max_score / (max_len - 1) : abort_score + 1;
        if (max_len > min_len && mean > abort_score)
        {
            int cur = pos;
            for (int j = max_str; j != -1; j = rec[j].p2anchor) { out_el[cur] = el[j]; out_score[cur] = rec[j].score2; cur++; }
            if (cur != pos)
            {
                if (n_chains > 0)
                    if ((float)(u64)(cur - pos + stale) / (float)(u64)(chain_off[1] - chain_off[0]) < stop_ratio) f_stop = 1;
                if (!f_stop && n_chains < max_chains)
                {
                    pos = cur;
                    n_chains++;
                    chain_off[n_chains] = pos;
                }
                else if (f_stop) stale += cur - pos;
            }
        }
    }
    return n_chains;
}

template <class E>
LNR_HD int traceback(const E * el, ChainRec * rec, int n, E * out_el, i32 * out_score, int * chain_off, int max_chains,
                     int min_len, int abort_score, int bestn, float stop_ratio)
{   // traceBackChains :307 -- number of distinct roots = records that start a tree
    int roots = 0;
    for (int i = 0; i < n; i++) roots += rec[i].p2anchor == -1;
    if (roots > 50) return traceback0(el, rec, n, out_el, out_score, chain_off, max_chains, min_len, abort_score, bestn, stop_ratio);
    return traceback1(el, rec, n, out_el, out_score, chain_off, max_chains, min_len, abort_score, bestn, stop_ratio);
}

// Warp versions of the two tracebacks: what is a scan over all n records in the reference -- the arg-max of every peel round
// (traceBackChains0), the search for the leaves (traceBackChains1), the count of roots -- runs on all lanes; the short,
// pointer-chasing parts (following one chain, ranking <= 64 trees) stay on lane 0. Same results as traceback<E> on lane 0.
template <class E>
LNR_PIPE int traceback0_w(const Warp & w, const E * el, ChainRec * rec, int n, E * out_el, i32 * out_score, int * chain_off, int max_chains,
                          int min_len, int abort_score, int bestn, float stop_ratio)
{   // traceBackChains0 :122
    const int delete_score = -1000;
    int n_chains = 0, pos = 0;                      // lane 0's
    if (w.lane == 0) chain_off[0] = 0;
    const int search_times = bestn < 50 ? bestn : 50;
    for (int it = 0; it < search_times; it++)
    {
        // the sequential scan `if (score > max) { max_2nd = max; max = score; .. }` from max = -1 ends with the FIRST index of
        // the largest score and with max_2nd = the largest score before that index (or -1)
        i32 best = -1; int bidx = 0x7fffffff;
        for (int j = w.lane; j < n; j += w.nl) { const i32 sc = rec[j].score; if (sc > best) { best = sc; bidx = j; } }
        const i32 max_score = wmax_i32(w, best);
        if (max_score <= -1) break;                 // f_done: nothing left (the stop-ratio clause cannot revive it: max_len = 0)
        const int max_str = -wmax_i32(w, best == max_score ? -bidx : (i32)0x80000000);
        if (max_score == 0) break;
        i32 pm = -1;
        for (int j = w.lane; j < max_str; j += w.nl) { const i32 sc = rec[j].score; pm = sc > pm ? sc : pm; }
        const i32 max_2nd = wmax_i32(w, pm);
        if (w.lane == 0)
        {
            const int max_len = rec[max_str].len;
            if (max_len > min_len && max_score / (max_len - 1) > abort_score)
            {
                int cur = pos;   // tentative chain [pos, cur)
                for (int j = max_str; j != -1; j = rec[j].p2anchor)
                {
                    if (rec[j].score != delete_score)
                    {
                        out_el[cur] = el[j];
                        out_score[cur] = rec[j].score2;
                        cur++;
                        rec[j].score = delete_score;
                    }
                    else
                    {
                        int infix = rec[j].score2;
                        if (max_score - infix < max_2nd)
                        {
                            for (int k = max_str; k != j; k = rec[k].p2anchor) rec[k].score = rec[k].score2 - infix;
                            cur = pos;
                        }
                        break;
                    }
                }
                if (cur != pos && n_chains < max_chains)
                {
                    pos = cur;
                    n_chains++;
                    chain_off[n_chains] = pos;
                }
            }
            rec[max_str].score = delete_score;
        }
        wsync(w);
    }
    (void)stop_ratio;
    return wbcast(w, n_chains, 0);
}
template <class E>
LNR_PIPE int traceback_w(const Warp & w, const E * el, ChainRec * rec, int n, E * out_el, i32 * out_score, int * chain_off, int max_chains,
                         int min_len, int abort_score, int bestn, float stop_ratio)
{   // traceBackChains :307
    int roots = 0;
    for (int i = w.lane; i < n; i += w.nl) roots += rec[i].p2anchor == -1;
    roots = wsum(w, roots);
    if (roots > 50) return traceback0_w(w, el, rec, n, out_el, out_score, chain_off, max_chains, min_len, abort_score, bestn, stop_ratio);
    // traceBackChains1 :214 -- only the leaves matter: gather them in index order (out_score is free until the chains are
    // written), then the reference's own loop over that short list
    int * leaf = (int *)out_score;
    int n_leaf = 0;
    for (int c = 0; c < n; c += w.nl)
    {
        const int j = c + w.lane;
        const bool is_leaf = j < n && rec[j].f_leaf != 0;
        const u32 bal = wballot(w, is_leaf);
        if (is_leaf) leaf[n_leaf + popc_below(w, bal)] = j;
        n_leaf += popc32(bal);
    }
    wsync(w);
    int n_chains = 0;
    if (w.lane == 0)
    {
        int root[64], lscore[64], llen[64], lidx[64];
        int nt = 0;
        for (int q = 0; q < n_leaf; q++)
        {
            const int j = leaf[q];
            int f_new = 1;
            for (int k = 0; k < nt; k++)
                if (root[k] == rec[j].root_ptr)
                {
                    if (rec[j].score > lscore[k]) { lscore[k] = rec[j].score; llen[k] = rec[j].len; lidx[k] = j; }
                    f_new = 0;
                }
            if (f_new && nt < 64) { root[nt] = rec[j].root_ptr; lscore[nt] = rec[j].score; llen[nt] = rec[j].len; lidx[nt] = j; nt++; }
        }
        struct Rank { int first, second; };
        Rank ranks[64];
        for (int i = 0; i < nt; i++) { ranks[i].first = i; ranks[i].second = lscore[i]; }
        gnu_sort(ranks, nt, [](const Rank & a, const Rank & b) { return a.second > b.second; });
        int pos = 0, f_stop = 0;
        int stale = 0;   // the reference keeps appending to an uncleared `chain` once f_stop is set (dead data)
        chain_off[0] = 0;
        int lim = bestn < nt ? bestn : nt;
        for (int i = 0; i < lim; i++)
        {
            int t = ranks[i].first;
            int max_score = lscore[t], max_len = llen[t], max_str = lidx[t];
            int mean = max_len > 1 ? max_score / (max_len - 1) : abort_score + 1;
            if (max_len > min_len && mean > abort_score)
            {
                int cur = pos;
                for (int j = max_str; j != -1; j = rec[j].p2anchor) { out_el[cur] = el[j]; out_score[cur] = rec[j].score2; cur++; }
                if (cur != pos)
                {
                    if (n_chains > 0)
                        if ((float)(u64)(cur - pos + stale) / (float)(u64)(chain_off[1] - chain_off[0]) < stop_ratio) f_stop = 1;
                    if (!f_stop && n_chains < max_chains)
                    {
                        pos = cur;
                        n_chains++;
                        chain_off[n_chains] = pos;
                    }
                    else if (f_stop) stale += cur - pos;
                }
            }
        }
    }
    wsync(w);
    return wbcast(w, n_chains, 0);
}

// ----------------------------------------------------------------------------------------------------
// blocks of hits / cords
// ----------------------------------------------------------------------------------------------------
// gather_blocks_ (pmpfinder.cpp:1484). Appends to sep[n_sep..]; when str_ends != null also records the
// shifted start/end cords. Returns the new n_sep.
LNR_HD int gather_blocks(u64 * cords, int n, YPair * str_ends, int & n_str_ends, Blk * sep, int n_sep, u32 str_, u32 end_,
                         u64 L, u64 large_gap, u64 cord_size, int f_set_end)
{
    n_str_ends = 0;
    if (n < 2) return n_sep;
    u64 dmax = cord_size / 2, d;
    u32 p_str = str_;
    for (u32 i = str_ + 1; i < end_; i++)
        if (is_end(cords[i - 1]) || !cords_consecutive(cords[i - 1], cords[i], large_gap))
        {
            if (str_ends)
            {
                d = umin64(L - cord_y(cords[p_str]) - 1, dmax);
                str_ends[n_str_ends].first = shift_cord(cords[p_str], (i64)d, (i64)d);
                d = umin64(L - cord_y(cords[i - 1]) - 1, dmax);
                str_ends[n_str_ends].second = shift_cord(cords[i - 1], (i64)d, (i64)d);
            }
            n_str_ends++;
            sep[n_sep].first = p_str; sep[n_sep].second = i; n_sep++;
            if (f_set_end) cords[i - 1] |= kFlagEnd;
            p_str = i;
        }
    if (str_ends)
    {
        d = umin64(L - cord_y(cords[n - 1]) - 1, dmax);
        str_ends[n_str_ends].first = shift_cord(cords[p_str], (i64)d, (i64)d);
        str_ends[n_str_ends].second = shift_cord(cords[n - 1], (i64)d, (i64)d);
    }
    n_str_ends++;
    sep[n_sep].first = p_str; sep[n_sep].second = (u32)n; n_sep++;
    return n_sep;
}

// preFilterChains2 (pmpfinder.cpp:2366) with getCordXY = get_cord_y. sep (nb blocks) is replaced by the
// y-disjoint pieces (tmp: scratch of capacity cap); cuts: 2*nb u64; strs: nb u64. Returns the new count,
// or -1 when the piece buffer would overflow.
LNR_HD int prefilter_chains2(u64 * hits, int n_hits, Blk * sep, int nb, Blk * tmp, int cap, u64 * cuts, u64 * strs)
{
    const u64 mask = 1ULL << 62;
    for (int i = 0; i < nb; i++)
    {
        cuts[2 * i] = sep[i].first;
        cuts[2 * i + 1] = (u64)(sep[i].second - 1) | mask;
        strs[i] = sep[i].first;
    }
    gnu_sort(cuts, 2 * nb, [hits, mask](const u64 & a, const u64 & b) { return cord_y(hits[a & ~mask]) < cord_y(hits[b & ~mask]); });
    // The reference scans every block linearly for every cut (pmpfinder.cpp:2398-2437). y is strictly ascending
    // inside a block (chains need dy >= 5), so the first k with y >= cuty is found by binary search, and a block
    // that ends below the cut cannot produce a piece -- same pieces, same order, without the quadratic rescans.
    int nt = 0;
    for (int i = 0; i < 2 * nb; i++)
    {
        const bool is_last = (cuts[i] & mask) != 0;
        u64 cuty = cord_y(hits[cuts[i] & ~mask]);
        for (int j = 0; j < nb && strs[j] < (u64)n_hits; j++)
        {
            if (strs[j] >= sep[j].second) continue;                       // block used up: the k loop is empty
            if (cuty < cord_y(hits[strs[j]])) continue;
            if (cord_y(hits[sep[j].second - 1]) < cuty) continue;          // no k with y >= cuty
            u64 lo = strs[j], hi = sep[j].second - 1;                     // first k in [lo, hi] with y >= cuty
            while (lo < hi)
            {
                u64 mid = (lo + hi) >> 1;
                if (cord_y(hits[mid]) >= cuty) hi = mid; else lo = mid + 1;
            }
            u64 up = (is_last && cord_y(hits[lo]) == cuty) ? lo + 1 : lo;
            if (strs[j] != up)
            {
                if (nt >= cap) return -1;
                tmp[nt].first = (u32)strs[j]; tmp[nt].second = (u32)up; nt++;
                strs[j] = up;
            }
        }
    }
    for (int i = 0; i < nt; i++) sep[i] = tmp[i];
    gnu_sort(sep, nt, [](const Blk & a, const Blk & b) { return a.second < b.second; });
    for (int i = 0; i < nt; i++) hits[sep[i].second - 1] |= kFlagEnd;
    return nt;
}

// getBestChains2 (cluster_util.cpp:469); mode 0 = hits blocks (getApxChainScore2), 1 = cord blocks
// (getApxChainScore3 on `strand`)
LNR_HD void best_chains2(const u64 * recs, const Blk * sep, const i32 * sep_score, ChainRec * ch, int nb, u64 L, int mode, int strand)
{
    const int depth = 20;
    for (int i = 0; i < nb; i++)
    {
        int j_str = i - depth > 0 ? i - depth : 0;
        int max_j = i, new_max = -1;
        for (int j = j_str; j < i; j++)
        {
            int s = mode == 0 ? score_blocks_hits(recs[sep[j].first], recs[sep[i].second - 1])
                              : score_blocks_cords(recs[sep[j].first], recs[sep[j].second - 1], recs[sep[i].first],
                                                   recs[sep[i].second - 1], L, strand);
            if (s > 0 && s + ch[j].score + sep_score[i] >= new_max)
            {
                max_j = j;
                new_max = s + ch[j].score + sep_score[i];
            }
        }
        if (new_max > 0)
        {
            ch[i].p2anchor = max_j; ch[i].score = new_max;
            ch[i].len = (i32)(sep[i].second - sep[i].first) + ch[max_j].len;
            ch[i].score2 = ch[i].score; ch[i].root_ptr = ch[max_j].root_ptr; ch[i].f_leaf = 1; ch[max_j].f_leaf = 0;
        }
        else
        {
            ch[i].p2anchor = -1; ch[i].score = sep_score[i]; ch[i].len = (i32)(sep[i].second - sep[i].first);
            ch[i].score2 = ch[i].score; ch[i].root_ptr = i; ch[i].f_leaf = 1;
        }
    }
}

// scratch needed by chain_blocks_base for nb blocks
struct BlockScratch
{
    u32 * ptr; Blk * sep_tmp; i32 * score_tmp; ChainRec * rec; Blk * out_el; i32 * out_score; int * chain_off;
};
template <class AllocT> LNR_PIPE_INL bool block_scratch_alloc(AllocT & ar, BlockScratch & s, int nb)
{
    s.ptr = arena_alloc<u32>(ar, nb);
    s.sep_tmp = arena_alloc<Blk>(ar, nb);
    s.score_tmp = arena_alloc<i32>(ar, nb);
    s.rec = arena_alloc<ChainRec>(ar, nb);
    s.out_el = arena_alloc<Blk>(ar, nb);
    s.out_score = arena_alloc<i32>(ar, nb);
    s.chain_off = arena_alloc<int>(ar, 8);
    return !ar.failed;
}

// chainBlocksBase (cluster_util.cpp:533). Returns the number of chains (<= 3, bestn), elements in s.out_el.
LNR_HD int chain_blocks_base(const u64 * recs, const Blk * sep, const i32 * sep_score, int nb, BlockScratch & s, u64 L,
                             int mode, int strand, int f_sort)
{
    if (nb < 2) return 0;
    for (int i = 0; i < nb; i++) s.ptr[i] = (u32)i;
    if (f_sort)
        gnu_sort(s.ptr, nb, [recs, sep](const u32 & a, const u32 & b) { return cord_x40(recs[sep[a].first]) > cord_x40(recs[sep[b].first]); });
    for (int i = 0; i < nb; i++) { s.sep_tmp[i] = sep[s.ptr[i]]; s.score_tmp[i] = sep_score[s.ptr[i]]; }
    best_chains2(recs, s.sep_tmp, s.score_tmp, s.rec, nb, L, mode, strand);
    return traceback<Blk>(s.sep_tmp, s.rec, nb, s.out_el, s.out_score, s.chain_off, 6, 1, 0, 3, 0.7f);
}

// ---- warp-cooperative statements of preFilterChains2 / getBestChains2 / chainBlocksBase for the hit blocks.
// The sequential ones above stay the reference statement (and serve the cord-block stage); these give the same
// results with the inner loops spread over the lanes: a read with several hundred blocks spent milliseconds in the
// blocks x cuts scan on one lane and set the kernel's tail.
LNR_PIPE int prefilter_chains2_w(const Warp & w, u32 * hist256, u64 * hits, int n_hits, Blk * sep, int nb, Blk * tmp, int cap, u64 * cuts,
                                 u64 * strs, u64 * s0, u64 * s1)
{
    const u64 mask = 1ULL << 62;
    for (int i = w.lane; i < nb; i += w.nl)
    {
        cuts[2 * i] = sep[i].first;
        cuts[2 * i + 1] = (u64)(sep[i].second - 1) | mask;
        strs[i] = sep[i].first;
    }
    wsync(w);
    // std::sort(cuts, by y of the hit) -- ties are common (block ends share y with the next block's start)
    {
        u64 * r = gnu_sort_w(w, hist256, cuts, s0, s1, 2 * nb, 20, sort_key(2, hits));
        if (r != cuts) { for (int i = w.lane; i < 2 * nb; i += w.nl) cuts[i] = r[i]; wsync(w); }
    }
    // per block: y of its current start (0xffffffff once it is used up) and y of its last hit, so that the cuts x blocks
    // scan is two coalesced loads and two compares per pair; only the rare pair that passes touches the hits
    u32 * ys = (u32 *)s1;
    u32 * ye = ys + nb;
    for (int j = w.lane; j < nb; j += w.nl)
    {
        ys[j] = sep[j].first < sep[j].second ? (u32)cord_y(hits[sep[j].first]) : 0xffffffffu;
        ye[j] = (u32)cord_y(hits[sep[j].second - 1]);
    }
    wsync(w);
    int nt = 0;
    int j_end = nb;      // the reference's block loop ends at the first block whose start pointer ran to the end of the hits
    for (int i = 0; i < 2 * nb; i++)
    {
        const u64 cut = cuts[i];
        const bool is_last = (cut & mask) != 0;
        const u32 cuty = (u32)cord_y(hits[cut & ~mask]);
        int j_end_next = j_end;
        for (int c = 0; c < j_end; c += w.nl)
        {
            const int j = c + w.lane;
            u64 up = 0, sj = 0;
            bool emit = false;
            if (j < j_end && ys[j] <= cuty && ye[j] >= cuty)
            {
                sj = strs[j];
                u64 lo = sj, hi = (u64)sep[j].second - 1;                 // first k in [lo, hi] with y >= cuty
                while (lo < hi)
                {
                    u64 mid = (lo + hi) >> 1;
                    if ((u32)cord_y(hits[mid]) >= cuty) hi = mid; else lo = mid + 1;
                }
                up = (is_last && (u32)cord_y(hits[lo]) == cuty) ? lo + 1 : lo;
                emit = sj != up;
            }
            const u32 be = wballot(w, emit);
            const int cnt = popc32(be);
            if (nt + cnt > cap) return -1;
            if (emit)
            {
                const int pos = nt + popc_below(w, be);
                tmp[pos].first = (u32)sj; tmp[pos].second = (u32)up;
                strs[j] = up;
                ys[j] = up < (u64)sep[j].second ? (u32)cord_y(hits[up]) : 0xffffffffu;
            }
            const u32 bend = wballot(w, emit && up >= (u64)n_hits);
            if (bend) { int je = c + ffs32(bend); if (je < j_end_next) j_end_next = je; }
            nt += cnt;
        }
        j_end = j_end_next;
        wsync(w);
    }
    {
        u64 * r = gnu_sort_w(w, hist256, (u64 *)tmp, s0, s1, nt, 32, sort_key(3));
        for (int i = w.lane; i < nt; i += w.nl) ((u64 *)sep)[i] = r[i];
        wsync(w);
    }
    for (int i = w.lane; i < nt; i += w.nl) wor_flag64(&hits[sep[i].second - 1], kFlagEnd);
    wsync(w);
    return nt;
}

// getBestChains2, mode 0. The sequential scan keeps the LAST j among equal sums (>=), so the lanes take j from
// i-1 downwards and a later chunk only wins with a strictly larger sum.
LNR_PIPE void best_chains2_hits_w(const Warp & w, const u64 * recs, const Blk * sep, const i32 * sep_score, ChainRec * ch, int nb)
{
    const int depth = 20;
    for (int i = 0; i < nb; i++)
    {
        const int j_str = i - depth > 0 ? i - depth : 0;
        const u64 ri = recs[sep[i].second - 1];
        const i32 si = sep_score[i];
        int max_j = i, new_max = -1;
        bool found = false;
        for (int c = 0; i - 1 - c >= j_str; c += w.nl)
        {
            const int j = i - 1 - c - w.lane;
            i32 val = (i32)0x80000000;
            if (j >= j_str)
            {
                int s = score_blocks_hits(recs[sep[j].first], ri);
                if (s > 0) val = s + ch[j].score + si;
            }
            const i32 vmax = wmax_i32(w, val);
            if (vmax >= -1 && (!found || vmax > new_max))
            {
                const u32 b = wballot(w, val == vmax);
                max_j = i - 1 - c - ffs32(b);
                new_max = vmax;
                found = true;
            }
        }
        if (w.lane == 0)
        {
            if (new_max > 0)
            {
                ch[i].p2anchor = max_j; ch[i].score = new_max;
                ch[i].len = (i32)(sep[i].second - sep[i].first) + ch[max_j].len;
                ch[i].score2 = ch[i].score; ch[i].root_ptr = ch[max_j].root_ptr; ch[i].f_leaf = 1; ch[max_j].f_leaf = 0;
            }
            else
            {
                ch[i].p2anchor = -1; ch[i].score = si; ch[i].len = (i32)(sep[i].second - sep[i].first);
                ch[i].score2 = ch[i].score; ch[i].root_ptr = i; ch[i].f_leaf = 1;
            }
        }
        wsync(w);
    }
}

// chainBlocksBase for the hit blocks (mode 0, sorted by x); the number of chains is returned on every lane
LNR_PIPE int chain_blocks_hits_w(const Warp & w, u32 * hist256, const u64 * recs, const Blk * sep, const i32 * sep_score, int nb, BlockScratch & s,
                                 u64 * s0, u64 * s1, u64 * s2)
{
    if (nb < 2) return 0;
    for (int i = w.lane; i < nb; i += w.nl) s2[i] = (u64)i;
    wsync(w);
    {
        u64 * r = gnu_sort_w(w, hist256, s2, s0, s1, nb, 40, sort_key(4, recs, sep));
        for (int i = w.lane; i < nb; i += w.nl) s.ptr[i] = (u32)r[i];
        wsync(w);
    }
    for (int i = w.lane; i < nb; i += w.nl) { s.sep_tmp[i] = sep[s.ptr[i]]; s.score_tmp[i] = sep_score[s.ptr[i]]; }
    wsync(w);
    best_chains2_hits_w(w, recs, s.sep_tmp, s.score_tmp, s.rec, nb);
    return traceback_w<Blk>(w, s.sep_tmp, s.rec, nb, s.out_el, s.out_score, s.chain_off, 6, 1, 0, 3, 0.7f);
}

// _filterBlocksHits (cluster_util.cpp:633): major chain + up to 4 optional chains > 0.8 * len.
// Plans the new hit list (without header): which chains survive. keep_chain[c] = 1 for the chains copied, in order.
// Returns the number of hits that will be written.
LNR_HD int filter_blocks_hits_plan(const Blk * el, const int * chain_off, int n_chains, u8 * keep_chain)
{
    u64 len_cur = 0;
    for (int i = chain_off[0]; i < chain_off[1]; i++) len_cur += el[i].second - el[i].first;
    int no = (int)len_cur;
    keep_chain[0] = 1;
    float major_bound = (float)(0.8 * (double)len_cur);
    u32 major_limit = 5, major_n = 1;
    for (int c = 1; c < n_chains; c++)
    {
        len_cur = 0;
        for (int i = chain_off[c]; i < chain_off[c + 1]; i++) len_cur += el[i].second - el[i].first;
        // (the reference's third branch needs len_cur == 0, impossible for non-empty blocks)
        keep_chain[c] = 0;
        if (major_n < major_limit && (float)len_cur > major_bound) { ++major_n; keep_chain[c] = 1; no += (int)len_cur; }
    }
    return no;
}
// copies the planned chains: every hit loses its end flag except the last of each chain. All lanes.
LNR_PIPE void filter_blocks_hits_copy(const Warp & w, const Blk * el, const int * chain_off, int n_chains, const u8 * keep_chain,
                                      const u64 * hits, u64 * out)
{
    int no = 0;
    for (int c = 0; c < n_chains; c++)
    {
        if (!keep_chain[c]) continue;
        for (int i = chain_off[c]; i < chain_off[c + 1]; i++)
        {
            int b = (int)el[i].first, e = (int)el[i].second;
            bool last_piece = i == chain_off[c + 1] - 1;
            for (int j = b + w.lane; j < e; j += w.nl)
            {
                u64 h = hits[j] & ~kFlagEnd;
                if (last_piece && j == e - 1) h |= kFlagEnd;
                out[no + (j - b)] = h;
            }
            no += e - b;
        }
    }
    wsync(w);
}

// ----------------------------------------------------------------------------------------------------
// window extension (pmpfinder.cpp:883-1176). Warp-uniform: every lane follows the same control flow and
// holds the same scalars; the 3 candidate windows x 2 scripts x 3 ints of one step are spread over 18
// lanes, cords are written by lane 0 and the most recent cord is kept in a register (`last`).
// ----------------------------------------------------------------------------------------------------
// distances of the read window at feature row y against genome windows x0, x0+1, x0+2 (__windowDist :655;
// the reference does not bounds-check, the guard only protects the device on out-of-spec data)
// What stays fixed along one walk (contig, strand), fetched once: the walk itself is a serial chain of window steps, so
// everything a step re-reads from the tables lengthens the chain.
struct WinCtx
{
    const F96 * fa;     // read features of the strand
    const F96 * fb;     // genome features of the contig
    const i16 * sa;     // -f 1: the same two strings as shorts
    const i16 * sb;
    int ft; u32 win;
    u32 nf1, nf2;
    u64 id, strand;
    // lane constants of the 18-way split of a step (full warp only): int offsets into the two rows and the sum slot
    int off_a, off_b, cand, shift;
    u32 windows;        // candidates evaluated, added to the counters by the caller
    u64 hi;             // contig and strand bits of every cord of the walk
};
template <int FT>
LNR_PIPE_INL WinCtx win_ctx(const Warp & w, const PipeIn & in, u64 cord)
{
    WinCtx c;
    c.id = cord_id(cord); c.strand = cord_strand(cord);
    c.ft = FT ? FT : in.ft; c.win = FT == 1 ? (u32)kWin32 : (FT == 2 ? (u32)kWin : in.win);   // FT: the feature type as a compile-time constant
    c.fa = nullptr; c.fb = nullptr; c.sa = nullptr; c.sb = nullptr;
    if (c.ft == 1) { c.sa = in.s1[c.strand]; c.sb = in.s2[c.id]; }
    else { c.fa = in.f1[c.strand]; c.fb = in.f2[c.id]; }
    c.nf1 = in.nf1; c.nf2 = in.nf2[c.id];
    int t = w.lane < 18 ? w.lane : 0;
    int cd = t / 6, part = t - 6 * cd, i = part >= 3 ? 3 : 0, k = part - i;
    c.cand = cd; c.off_a = 3 * i + k; c.off_b = 3 * (cd + i) + k; c.shift = 10 * cd;
    c.windows = 0;
    c.hi = (c.id << 50) + (c.strand << 61);
    return c;
}
LNR_PIPE_INL void wdist3(const Warp & w, WinCtx & c, u32 y, u32 x0, u32 d[3])
{
    c.windows += 3;
    const u32 nf2 = c.nf2;
    const bool yok = y + 3 < c.nf1;
    const F96 * fa = c.fa + y;
    const F96 * fb = c.fb + x0;
    // every script distance is <= 5*32, a window sums 6 of them: 10 bits per candidate, one warp add
    int packed = 0;
    if (w.nl == 32)
    {
        if (w.lane < 18 && yok && x0 + c.cand + 3 < nf2)
            packed = script_dist(((const i32 *)fa)[c.off_a], ((const i32 *)fb)[c.off_b]) << c.shift;
    }
    else
    {
        for (int t = w.lane; t < 18; t += w.nl)
        {
            int cd = t / 6, part = t - 6 * cd, i = part >= 3 ? 3 : 0, k = part - i;
            if (yok && x0 + cd + 3 < nf2) packed += script_dist(fa[i].v[k], fb[cd + i].v[k]) << (10 * cd);
        }
    }
    packed = wsum(w, packed);
    int s0 = packed & 1023, s1 = (packed >> 10) & 1023, s2 = (packed >> 20) & 1023;
    d[0] = (yok && x0 + 3 < nf2) ? (u32)s0 : 1000u;
    d[1] = (yok && x0 + 4 < nf2) ? (u32)s1 : 1000u;
    d[2] = (yok && x0 + 5 < nf2) ? (u32)s2 : 1000u;
}
// -f 1: minimum over the 6 candidate genome windows x0 .. x0+5 (first strict minimum in ascending x); one candidate per lane
LNR_PIPE_INL void wmin6_32(const Warp & w, WinCtx & c, u32 y, u32 x0, u32 & mn, u32 & x_min)
{
    c.windows += 6;
    if (w.nl == 32)
    {
        u32 key = 0xffffffffu;
        if (w.lane < 6) key = (window_dist32(c.sa, c.nf1, y, c.sb, c.nf2, x0 + (u32)w.lane) << 3) | (u32)w.lane;   // distance < 2^10
        key = ~wmax_u32(w, ~key);
        mn = key >> 3; x_min = x0 + (key & 7u);
    }
    else
    {
        mn = 0xffffffffu; x_min = x0;
        for (u32 q = 0; q < 6; q++)
        {
            u32 d = window_dist32(c.sa, c.nf1, y, c.sb, c.nf2, x0 + q);
            if (d < mn) { mn = d; x_min = x0 + q; }
        }
    }
}
LNR_PIPE_INL bool previous_step32(const Warp & w, WinCtx & c, u32 & xr, u32 & yr)   // previousWindow :883 with ApxMapParm1_32
{
    const u32 x_suf = xr, y_suf = yr;
    if (y_suf < (u32)kMed32 || x_suf < (u32)kSup32) return false;
    const u32 y = y_suf - kMed32;
    u32 mn, x_min;
    wmin6_32(w, c, y, x_suf - kSup32, mn, x_min);
    if (mn > (u32)kWinThr) return false;
    if (x_suf - x_min > (u32)kMed32) { xr = x_suf - kMed32; yr = x_suf - x_min - kMed32 + y; }
    else { xr = x_min; yr = y; }
    return c.hi != 0 || xr != 0 || yr != 0;
}
LNR_PIPE_INL bool next_step32(const Warp & w, WinCtx & c, u32 & xr, u32 & yr)       // nextWindow :1079 with ApxMapParm1_32
{
    const u32 x_pre = xr, y_pre = yr;
    if (y_pre + 2 * kSup32 > c.nf1 || x_pre + 2 * kSup32 > c.nf2) return false;
    const u32 y = y_pre + kMed32;
    u32 mn, x_min;
    wmin6_32(w, c, y, x_pre + kInf32, mn, x_min);
    if (mn > (u32)kWinThr) return false;
    if (x_min - x_pre > (u32)kMed32) { xr = x_pre + kMed32; yr = x_pre + kMed32 - x_min + y; }
    else { xr = x_min; yr = y; }
    return true;
}
// One window step in feature-row coordinates (x row = cord_x >> 4, y row = cord_y >> 4; every cord a step produces has
// zero low nibbles, so the rows carry the whole state of a walk): 32-bit arithmetic, the 64-bit cord is assembled from
// the walk's constant contig/strand bits only when it is stored. Returns false where the reference returns 0.
LNR_PIPE_INL bool previous_step(const Warp & w, WinCtx & c, u32 & xr, u32 & yr)   // previousWindow :883
{
    if (c.ft == 1) return previous_step32(w, c, xr, yr);
    const u32 x_suf = xr, y_suf = yr;
    if (y_suf < (u32)kMed || x_suf < (u32)kSup) return false;
    const u32 y = y_suf - kMed, x0 = x_suf - kSup;  // candidates x_suf-6 .. x_suf-4
    u32 d[3];
    wdist3(w, c, y, x0, d);
    u32 mn = d[0], x_min = x0;                      // first strict minimum in ascending x
    if (d[1] < mn) { mn = d[1]; x_min = x0 + 1; }
    if (d[2] < mn) { mn = d[2]; x_min = x0 + 2; }
    if (mn > (u32)kWinThr) return false;
    if (x_suf - x_min > (u32)kMed) { xr = x_suf - kMed; yr = x_suf - x_min - kMed + y; }
    else { xr = x_min; yr = y; }
    return c.hi != 0 || xr != 0 || yr != 0;         // a cord that is numerically 0 reads as "no window" (:1158)
}
LNR_PIPE_INL bool next_step(const Warp & w, WinCtx & c, u32 & xr, u32 & yr)   // nextWindow :1079
{
    if (c.ft == 1) return next_step32(w, c, xr, yr);
    const u32 x_pre = xr, y_pre = yr;
    if (y_pre + 2 * kSup > c.nf1 || x_pre + 2 * kSup > c.nf2) return false;
    const u32 y = y_pre + kMed, x0 = x_pre + kInf;  // candidates x_pre+3 .. x_pre+5
    u32 d[3];
    wdist3(w, c, y, x0, d);
    u32 mn = d[0], x_min = x0;
    if (d[1] < mn) { mn = d[1]; x_min = x0 + 1; }
    if (d[2] < mn) { mn = d[2]; x_min = x0 + 2; }
    if (mn > (u32)kWinThr) return false;
    if (x_min - x_pre > (u32)kMed) { xr = x_pre + kMed; yr = x_pre + kMed - x_min + y; }
    else { xr = x_min; yr = y; }
    return true;                                     // y >= kMed: never the zero cord
}
LNR_PIPE_INL u64 win_cord(const WinCtx & c, u32 xr, u32 yr) { return c.hi + ((u64)xr << 24) + ((u64)yr << 4); }
// extendWindow :1152; returns false when the cord buffer is full. `last` == cords[n-1] on entry and exit.
template <int FT>
LNR_PIPE_INL bool extend_window(const Warp & w, const PipeIn & in, u64 * cords, int & n, int cap, u64 & last, u64 ystr, u64 yend,
                                PipeCounters & cnt)
{
    int p_str = n - 1;
    const u64 first = last;
    // every cord of the walk carries the contig and strand of its seeding cord (the windows only move x and y)
    WinCtx wc = win_ctx<FT>(w, in, last);
    const u32 x_first = (u32)(cord_x(last) >> 4), y_first = (u32)(cord_y(last) >> 4);
    u32 xr = x_first, yr = y_first;
    while (previous_step(w, wc, xr, yr) && ((u64)yr << 4) >= ystr)
    {
        if (n >= cap) { cnt.windows += wc.windows; return false; }
        last = win_cord(wc, xr, yr);
        if (w.lane == 0) cords[n] = last;
        n++;
    }
    if (n - p_str > 1)
    {
        wsync(w);
        for (int k = p_str + w.lane; k < (p_str + n) / 2; k += w.nl)
        {
            u64 t = cords[k]; cords[k] = cords[n - k + p_str - 1]; cords[n - k + p_str - 1] = t;
        }
        wsync(w);
    }
    last = first;                                   // after the reversal the seeding cord is last again
    xr = x_first; yr = y_first;
    while (next_step(w, wc, xr, yr) && ((u64)yr << 4) + wc.win < yend)
    {
        if (n >= cap) { cnt.windows += wc.windows; return false; }
        last = win_cord(wc, xr, yr);
        if (w.lane == 0) cords[n] = last;
        n++;
    }
    cnt.windows += wc.windows;
    return true;
}

// path_dst_2 (pmpfinder.cpp:1309), iterators restated as indices (hitBegin = 1). Warp-uniform.
// Returns false on overflow.
// FT = 0: the feature type is read from `in` (re-map / big-arena passes, host build); 1 / 2: compiled for that type alone, so
// that the 2_48 walk of the extension kernel carries none of the 1_32 code
template <int FT = 0>
LNR_PIPE bool path_dst_2(const Warp & w, const PipeIn & in, const u64 * H, int nh, u64 * cords, int & nc, int cap, u64 read_str,
                         u64 read_end, PipeCounters & cnt)
{
    const u64 L = in.L, cs = FT == 1 ? (u64)kWin32 : (FT == 2 ? (u64)kWin : (u64)in.win);
    int hb = 1, he = nh;
    if (hb + 1 >= he) return true;
    u64 last;
    if (nc == 0)                                    // initCords
    {
        if (cap < 1) return false;
        if (w.lane == 0) cords[0] = kFlagEnd;
        nc = 1;
        last = kFlagEnd;
    }
    else last = cords[nc - 1];
    u64 ready_str, ready_end, cordy_str = 0, cordy_end = 0;
    bool f_sp_l = false, f_sp_r = false, f_block_end = false, f_append = false;
    int nx = hb + 1, first = hb;
    for (int it = hb; it < he; it = nx++)
    {
        const u64 Hit = H[it], Hprev = H[it - 1];
        ready_str = cord_strand(Hit) ? L - read_end : read_str;
        ready_end = cord_strand(Hit) ? L - read_str + 1 : read_end;
        bool it_first = is_end(Hprev);
        i64 da_l = it_first ? 0 : iabs64((i64)(cord_x(Hit) - cord_x(Hprev) - cord_y(Hit) + cord_y(Hprev)));
        f_sp_l = (da_l > 80) || cord_strand(Hit ^ Hprev);
        u64 Hn1 = Hit;                              // H[nx - 1]
        while (1)
        {
            if (nx >= he || is_end(Hn1)) { f_block_end = true; first = nx; break; }
            u64 Hn = H[nx];
            i64 da_r = iabs64((i64)(cord_x(Hn) - cord_x(Hn1) - cord_y(Hn) + cord_y(Hn1)));
            f_sp_r = (da_r > 80) || cord_strand(Hn ^ Hn1);
            if ((cord_y(Hit) + cs < cord_y(Hn) && cord_x(Hit) + cs < cord_x(Hn)) || f_sp_r) break;
            nx++;
            Hn1 = Hn;
        }
        if (!f_sp_r && !f_block_end)
        {
            // sic: the whole hit value, not its y, when f_sp_l (pmpfinder.cpp:1360)
            cordy_str = f_sp_l ? Hit : (it_first ? ready_str : cord_y(last));
            cordy_end = cord_y(H[nx]);
            if (nc >= cap) return false;
            last = Hit & ~kFlagEnd;
            if (w.lane == 0) cords[nc] = last;
            nc++;
            f_append = true;
        }
        else
        {
            if (!f_sp_l && cord_y(Hn1) >= cs && cord_x(Hn1) >= cs)
            {
                u64 ncord = shift_cord(Hn1, -(i64)cs, -(i64)cs);
                cordy_str = it_first ? read_str : cord_y(ncord);
                cordy_end = cord_y(Hn1);
                if (nc >= cap) return false;
                last = ncord & ~kFlagEnd;
                if (w.lane == 0) cords[nc] = last;
                nc++;
                f_append = true;
            }
            else f_append = false;
        }
        if (is_end(Hit) || f_block_end) { f_block_end = true; cordy_end = ready_end; }
        if (f_append)
            if (!extend_window<FT>(w, in, cords, nc, cap, last, cordy_str, cordy_end, cnt)) return false;
        if (f_block_end)
        {
            last |= kFlagEnd;
            if (w.lane == 0) cords[nc - 1] = last;
        }
        nx = f_block_end ? first : nx;
        f_sp_l = f_sp_r = f_block_end = f_append = false;
    }
    wsync(w);
    return true;
}

// clean_blocks_ (pmpfinder.cpp:1537); returns the new length
LNR_HD int clean_blocks(u64 * cords, int n, u64 drop_len, i64 map_err)
{
    if (n == 0) return 0;
    u64 ptr = 1, len = 0;
    for (int i = 1; i < n; i++)
    {
        len++;
        if (!is_end(cords[i - 1]))
        {
            i64 dx = (i64)(cord_x(cords[i]) - cord_x(cords[ptr - 1]));
            i64 dy = (i64)(cord_y(cords[i]) - cord_y(cords[ptr - 1]));
            if (dx < 0 || dy < 0)
            {
                if (iabs64(dx) < map_err && iabs64(dy) < map_err) { --len; --ptr; }
                else cords[ptr] = cords[i];
            }
            else cords[ptr] = cords[i];
        }
        else cords[ptr] = cords[i];
        if (is_end(cords[i]))
        {
            ptr = len < drop_len ? ptr - len : ptr;
            len = 0;
            cords[ptr] |= kFlagEnd;
        }
        ptr++;
    }
    return (int)ptr;
}

LNR_HD void forward_y(YPair se, u64 L, u64 & a, u64 & b)   // getUPForwardy cords.cpp:469
{
    if (cord_strand(se.first)) { a = L - cord_y(se.second) - 1; b = L - cord_y(se.first) - 1; }
    else { a = cord_y(se.first); b = cord_y(se.second); }
}
LNR_HD int push_gap(YPair * gaps, int & ng, int cap, u64 a, u64 b, u64 L)
{
    if (ng >= cap) return 0;
    gaps[ng].first = a; gaps[ng].second = b;
    u64 ga, gb;
    forward_y(gaps[ng], L, ga, gb);
    ng++;
    return (int)(gb - ga);
}
// gather_gaps_y_ (pmpfinder.cpp:1592). Returns the summed gap length; gaps appended to gaps[0..ng)
LNR_HD int gather_gaps_y(YPair * str_ends, int ns, YPair * gaps, int & ng, int cap, u64 L, u64 gap_size)
{
    u64 cord_frt = 0, cord_end = L - 1;
    int sum = 0;
    ng = 0;
    if (ns == 0) return push_gap(gaps, ng, cap, cord_frt, cord_end, L);
    gnu_sort(str_ends, ns, [L](const YPair & i, const YPair & j) {
        u64 y1 = cord_strand(i.first) ? L - cord_y(i.second) - 1 : cord_y(i.first);
        u64 y2 = cord_strand(j.first) ? L - cord_y(j.second) - 1 : cord_y(j.first);
        return y1 < y2;
    });
    u64 f_cover = 0, cordy1 = 0, cordy2 = 0, y1a, y1b, y2a, y2b;
    forward_y(str_ends[0], L, y1a, y1b);
    y2a = y1a; y2b = y1b;
    if (y1a > gap_size)
    {
        cordy2 = cord_y(y1a);
        sum += push_gap(gaps, ng, cap, cord_frt, cordy2, L);
    }
    for (int i = 1; i < ns; i++)
    {
        if (!f_cover) { forward_y(str_ends[i - 1], L, y1a, y1b); cordy1 = cord_y(y1b); }
        forward_y(str_ends[i], L, y2a, y2b);
        cordy2 = cord_y(y2a);
        if (y1b > y2b) f_cover = 1;
        else
        {
            if (y2a > y1b && y2a - y1b > gap_size) sum += push_gap(gaps, ng, cap, cordy1, cordy2, L);
            f_cover = 0;
        }
    }
    u64 max_y_end = f_cover ? y1b : y2b;
    if (L - max_y_end > gap_size) sum += push_gap(gaps, ng, cap, max_y_end, cord_end, L);
    return sum;
}

// ----------------------------------------------------------------------------------------------------
// cord-block chaining (cluster_util.cpp:936-1103)
// ----------------------------------------------------------------------------------------------------
LNR_HD void revert_chain_block_strand(Blk * el, int * chain_off, int n_chains, const u64 * cords, int strand)
{   // revertChainBlockStrand :1023 (the reference appends a (0,0) terminator; restated with an index bound)
    u64 f_strand = strand ? 1 : 0;
    for (int c = 0; c < n_chains; c++)
    {
        Blk * ch = el + chain_off[c];
        int len = chain_off[c + 1] - chain_off[c];
        u64 pre = 0, cur = 0;
        int swap_str = 0;
        for (int j = 0; j <= len; j++)
        {
            if (j == len || cord_strand(cords[ch[j].first]) == f_strand) cur = 0;
            else cur = 1;
            if (cur && !pre) swap_str = j;
            if (!cur && pre)
                for (int k = swap_str; k < (swap_str + j) / 2; k++)
                {
                    Blk t = ch[k]; ch[k] = ch[swap_str + j - 1 - k]; ch[swap_str + j - 1 - k] = t;
                }
            pre = cur;
        }
    }
}
// _filterBlocksCords :865 (f_header = 1, <= major_limit chains). Returns the new length written to out.
LNR_HD int filter_blocks_cords(const Blk * el, const int * chain_off, int n_chains, const u64 * cords, u64 * out, u32 major_limit)
{
    int no = 0;
    out[no++] = cords[0];
    u64 len_cur = 0;
    for (int i = chain_off[0]; i < chain_off[1]; i++)
    {
        for (u32 j = el[i].first; j < el[i].second; j++) out[no++] = cords[j] & ~kFlagEnd;
        len_cur += el[i].second - el[i].first;
    }
    out[no - 1] |= kFlagEnd;
    float major_bound = (float)(0.8 * (double)len_cur);
    u32 major_n = 1;
    for (int c = 1; c < n_chains && major_n < major_limit; c++)
    {
        len_cur = 0;
        for (int i = chain_off[c]; i < chain_off[c + 1]; i++) len_cur += el[i].second - el[i].first;
        if ((float)len_cur > major_bound)
        {
            ++major_n;
            for (int i = chain_off[c]; i < chain_off[c + 1]; i++)
                for (u32 k = el[i].first; k < el[i].second; k++) out[no++] = cords[k] & ~kFlagEnd;
            out[no - 1] |= kFlagEnd;
        }
    }
    return no;
}

LNR_HD u64 block_key_y(const u64 * cords, Blk a, u64 L, int strand)
{
    bool flip = strand ? !cord_strand(cords[a.first]) : (cord_strand(cords[a.first]) != 0);
    return flip ? L - 1 - cord_y(cords[a.second - 1]) : cord_y(cords[a.first]);
}
// chainBlocksSingleStrand :936
LNR_HD int chain_blocks_single_strand(const u64 * cords, Blk * sep, int nb, BlockScratch & s, i32 * sep_score, int strand, u64 L,
                                      u32 init_score)
{
    gnu_sort(sep, nb, [cords, L, strand](const Blk & a, const Blk & b) { return block_key_y(cords, a, L, strand) > block_key_y(cords, b, L, strand); });
    for (int i = 0; i < nb; i++) sep_score[i] = (i32)((sep[i].second - sep[i].first) * init_score);
    return chain_blocks_base(cords, sep, sep_score, nb, s, L, 1, strand, 0);
}
LNR_HD int best_strand_of(const Blk * e1, const int * o1, int n1, const Blk * e2, const int * o2, int n2)
{   // getChainBlocksBestStrand :979
    int l1 = 0, l2 = 0;
    int m = n1 < n2 ? n1 : n2;
    for (int i = 0; i < m; i++)
    {
        for (int j = o1[i]; j < o1[i + 1]; j++) l1 += (int)(e1[j].second - e1[j].first);
        for (int j = o2[i]; j < o2[i + 1]; j++) l2 += (int)(e2[j].second - e2[j].first);
        if (l1 < l2) return 1;
        else if (l1 > l2) return 0;
    }
    return 0;
}

// ----------------------------------------------------------------------------------------------------
// warp-parallel versions of the two linear passes over the cord list
// ----------------------------------------------------------------------------------------------------
// clean_blocks_ is the identity when no cord steps back inside a block and no block is shorter than
// drop_len; that is the common case and is checked by all lanes. Otherwise lane 0 runs clean_blocks.
LNR_PIPE int clean_blocks_w(const Warp & w, u64 * cords, int n, u64 drop_len, i64 map_err)
{
    if (n == 0) return 0;
    wsync(w);
    bool bad = false;
    for (int i = 1 + w.lane; i < n; i += w.nl)
    {
        u64 p = cords[i - 1], c = cords[i];
        if (!is_end(p))
        {
            i64 dx = (i64)(cord_x(c) - cord_x(p)), dy = (i64)(cord_y(c) - cord_y(p));
            if (dx < 0 || dy < 0) bad = true;
        }
        // a block [s, i] shorter than drop_len (<= 2): its end cord directly or one step after another end
        if (is_end(c) && drop_len > 0)
        {
            if (is_end(p) && 1 < drop_len) bad = true;
            if (i >= 2 && !is_end(p) && is_end(cords[i - 2]) && 2 < drop_len) bad = true;
        }
    }
    if (wballot(w, bad) == 0) return n;
    int out = n;
    if (w.lane == 0) out = clean_blocks(cords, n, drop_len, map_err);
    out = wbcast(w, out, 0);
    wsync(w);
    return out;
}

// gather_blocks_ (pmpfinder.cpp:1484), all lanes: block boundaries are found 32 cords at a time and appended in
// order. Same outputs as gather_blocks (str_ = 1, end_ = n).
LNR_PIPE int gather_blocks_w(const Warp & w, u64 * cords, int n, YPair * str_ends, int & n_str_ends, Blk * sep, u64 L, u64 large_gap,
                             u64 cord_size, int f_set_end)
{
    n_str_ends = 0;
    if (n < 2) return 0;
    wsync(w);
    const u64 dmax = cord_size / 2;
    int nb = 0;
    u32 p_str = 1;
    for (int c0 = 2; c0 < n; c0 += w.nl)
    {
        int i = c0 + w.lane;
        bool brk = false;
        u64 prev = 0;
        if (i < n)
        {
            prev = cords[i - 1];
            brk = is_end(prev) || !cords_consecutive(prev, cords[i], large_gap);
        }
        u32 m = wballot(w, brk);
        int rank = popc_below(w, m);
        // start of this lane's block = previous boundary (in this chunk or carried)
        u32 below = w.nl == 32 ? (m & ((1u << w.lane) - 1)) : 0;
        u32 start = below ? (u32)(c0 + (hibit32(below))) : p_str;
        if (brk)
        {
            if (str_ends)
            {
                u64 sc = cords[start];
                u64 d = umin64(L - cord_y(sc) - 1, dmax);
                str_ends[nb + rank].first = shift_cord(sc, (i64)d, (i64)d);
                d = umin64(L - cord_y(prev) - 1, dmax);
                str_ends[nb + rank].second = shift_cord(prev, (i64)d, (i64)d);
            }
            sep[nb + rank].first = start; sep[nb + rank].second = (u32)i;
            if (f_set_end) cords[i - 1] = prev | kFlagEnd;
        }
        if (m) p_str = (u32)(c0 + (hibit32(m)));
        nb += popc32(m);
    }
    if (w.lane == 0)
    {
        if (str_ends)
        {
            u64 d = umin64(L - cord_y(cords[n - 1]) - 1, dmax);
            str_ends[nb].first = shift_cord(cords[p_str], (i64)d, (i64)d);
            str_ends[nb].second = shift_cord(cords[n - 1], (i64)d, (i64)d);
        }
        sep[nb].first = p_str; sep[nb].second = (u32)n;
    }
    nb++;
    n_str_ends = nb;
    wsync(w);
    return nb;
}

// ----------------------------------------------------------------------------------------------------
// Phase 1: one apxMap_ call after seeding (pmpfinder.cpp:2632): filter -> chain -> hits -> blocks ->
// window extension. A[0..n): raw anchors with A[0] the sentinel slot; B: second buffer of n entries.
// Appends to cords. dbg_hits (optional): hits after getAnchorHitsChains.
// Return: 0 ok, 1 scratch/cord capacity exhausted.
// ----------------------------------------------------------------------------------------------------
// upper bound of the arena bytes phase_map claims for n raw anchors (incl. sentinel): every array is <= n entries
LNR_HD u64 phase_map_scratch_bound(int n)
{
    u64 m = (u64)n + 2;
    return m * (8 + 4 + 8 + 8 + 4 + 24 + 8 + 8 + 8 + 16 + 8 + 4 + 52 + 1 + 40) + 2048;
}

// The hit stage of apxMap_ in three sections. phase_map runs them back to back inside one warp (re-map pass, big-arena
// pass, host emulation); the primary pass runs each section as its own kernel (k_hits_sort / k_hits_chain /
// k_hits_blocks) so that every warp of an SM executes the same few KB of code and each kernel gets its own register
// budget -- as one 250 KB kernel the warps of an SM sat in different sections and starved on instruction fetch.

// Section 1: filterAnchors (:2159: binningFilter + filterAnchors1) and the AnchorX sort of chainAnchorsHits (:2465).
// hist256: the warp's kWarpSmemWords words of shared memory (radix histogram + zeroed bin sketch).
// A, B: the task's two n-entry regions (A = raw anchors, A[0] the sentinel slot). On return X[0..n2) are the anchors in
// chaining order; X is A, B or arena memory. rc: 0 = continue with section 2 (n2 >= 2), 3 = the read has no hits,
// 1 = arena exhausted.
LNR_PIPE int hits_sec_sort(const Warp & w, Arena & ar, u32 * hist256, u32 * bins, u64 * A, u64 * B, int n, PipeCounters & cnt,
                           long long & tl, u64 *& X, int & n2)
{
    X = A; n2 = 0;
    if (w.lane == 0) A[0] = 0;               // Anchors::init(1) base.cpp:272
    wsync(w);
    int m;
    u64 * S = binning_filter(w, bins, hist256 + 256, A, B, n, m);
    LNR_LAP(cnt, 0, tl);
    if (m <= 1) return 3;                    // no chains, hits stay empty (path_dst :1457)
    u64 * O = (S == A) ? B : A;              // the other buffer
    if (w.lane == 0) S[0] = 0;               // filterAnchorsList :2030 overwrites whatever is first
    wsync(w);
    u64 * C = arena_alloc<u64>(ar, (u64)m);
    Blk * ranges = arena_alloc<Blk>(ar, (u64)m / 2 + 2);
    if (ar.failed) return 1;
    u64 * sorted = radix_sort(w, hist256, S, O, C, m, 62, sort_key(0));
    LNR_LAP(cnt, 1, tl);
    const int nr = filter_anchor_runs_w(w, sorted, m, ranges);
    // compact the accepted runs (anchors[0] is dropped, filterAnchors1 :2073)
    u64 * F = (sorted == O) ? C : O;         // a buffer that is neither `sorted` nor S
    if (F == S) F = (sorted == C) ? O : C;
    for (int r = 0; r < nr; r++)
    {
        int b = (int)ranges[r].first, e = (int)ranges[r].second;
        for (int j = b + w.lane; j < e; j += w.nl) F[n2 + (j - b)] = sorted[j];
        n2 += e - b;
    }
    wsync(w);
    LNR_LAP(cnt, 2, tl);
    X = F;
    // a single anchor is sorted trivially and chainAnchorsBase returns without chains (cluster_util.cpp:450): with
    // fewer than two anchors the hit list stays [sentinel] and path_dst returns at :1457
    if (n2 < 2) return 3;
    // ---- chainAnchorsHits (:2448): sort by AnchorX descending with std::sort's tie order
    u64 * T0 = (F == A) ? B : A;             // free buffers: the two of {A,B,C} that are not F
    u64 * T1 = (F == C) ? B : C;
    if (T1 == T0) T1 = C;
    X = radix_sort(w, hist256, F, T0, T1, n2, 30, sort_key(1));
    int tie = 0;
    for (int i = 1 + w.lane; i < n2; i += w.nl) tie |= anchor_x(X[i]) == anchor_x(X[i - 1]);
    tie = wballot(w, tie != 0) != 0;
    if (tie)
    {
        // comparator ties: only libstdc++'s own permutation is right
        X = gnu_sort_w(w, hist256, F, T0, T1, n2, 30, sort_key(1));
        cnt.t[11]++;
    }
    LNR_LAP(cnt, 3, tl);
    return 0;
}

// Section 2: chainAnchorsBase DP + traceback (getBestChains / traceBackChains, cluster_util.cpp:450-560) over
// X[0..n2), n2 >= 2. Writes hits[0..n_hits) (hits[0] the sentinel) and their scores. hits / hits_score: n2 + 2 entries.
LNR_PIPE int hits_sec_chain(const Warp & w, Arena & ar, const PipeIn & in, const u64 * X, int n2, int score_type, u64 * hits,
                            i32 * hits_score, int & n_hits, PipeCounters & cnt, long long & tl)
{
    ChainRec * rec = arena_alloc<ChainRec>(ar, (u64)n2);
    u64 * ch_el = arena_alloc<u64>(ar, (u64)n2);
    int * chain_off = arena_alloc<int>(ar, 64);
    if (ar.failed) return 1;
    n_hits = 1;
    if (w.lane == 0) { hits[0] = kFlagEnd; hits_score[0] = 0; }
    best_chains(w, X, rec, n2, score_type);
    LNR_LAP(cnt, 4, tl);
    {
        const int nch = traceback_w<u64>(w, X, rec, n2, ch_el, hits_score + 1, chain_off, 60, 1, 45, 50, in.stop_ratio);
        for (int c = 0; c < nch; c++)
        {
            for (int j = chain_off[c] + w.lane; j < chain_off[c + 1]; j += w.nl)
                hits[1 + j] = hit2cord_dstr(ch_el[j]) | (j == chain_off[c + 1] - 1 ? kFlagEnd : 0ULL);
        }
        n_hits = 1 + chain_off[nch];
    }
    wsync(w);
    LNR_LAP(cnt, 5, tl);
    return 0;
}

// Section 3: blocks of hits -- gather_blocks_ (:1484) -> preFilterChains2 (:2366) -> chainBlocksHits -> _filterHits
// (:1417). hits[0..n_hits) with scores in; the surviving hits are written to out[0..n_hits) (out must not alias hits;
// capacity: the incoming n_hits). rc 0 = hits in out, 3 = fewer than two hits (the read has none), 1 = arena exhausted.
LNR_PIPE int hits_sec_blocks(const Warp & w, Arena & ar, u32 * hist256, const PipeIn & in, u64 * hits, const i32 * hits_score, int & n_hits,
                             u64 * out, u64 * dbg_hits, u32 * dbg_nhits, u32 dbg_hits_cap, PipeCounters & cnt, long long & tl)
{
    u64 * H = hits;
    Blk * sep = arena_alloc<Blk>(ar, (u64)n_hits + 1);
    Blk * sep_tmp = arena_alloc<Blk>(ar, (u64)n_hits + 1);
    u64 * cuts = arena_alloc<u64>(ar, 2 * (u64)n_hits + 2);
    u64 * strs = arena_alloc<u64>(ar, (u64)n_hits + 1);
    i32 * sep_score = arena_alloc<i32>(ar, (u64)n_hits + 1);
    u64 * hits2 = arena_alloc<u64>(ar, (u64)n_hits + 2);
    BlockScratch bs;
    block_scratch_alloc(ar, bs, n_hits + 1);
    u8 * keep_chain = arena_alloc<u8>(ar, 16);
    u64 * srt0 = arena_alloc<u64>(ar, 2 * (u64)n_hits + 2);     // scratch of the std::sort emulation (cuts: 2 per block)
    u64 * srt1 = arena_alloc<u64>(ar, 2 * (u64)n_hits + 2);
    u64 * srt2 = arena_alloc<u64>(ar, (u64)n_hits + 2);
    if (ar.failed) return 1;
    int nb0;
    {
        int dummy = 0;
        nb0 = gather_blocks_w(w, hits, n_hits, (YPair *)0, dummy, sep, in.L, 600, 0, 0);
    }
    int nb = prefilter_chains2_w(w, hist256, hits, n_hits, sep, nb0, sep_tmp, n_hits + 1, cuts, strs, srt0, srt1);
    if (nb < 0) return 1;
    for (int i = w.lane; i < nb; i += w.nl) sep_score[i] = hits_score[sep[i].first] - hits_score[sep[i].second - 1];
    wsync(w);
    int nch = chain_blocks_hits_w(w, hist256, hits, sep, sep_score, nb, bs, srt0, srt1, srt2);
    if (nch > 0)
    {
        if (w.lane == 0) n_hits = filter_blocks_hits_plan(bs.out_el, bs.chain_off, nch, keep_chain);
        n_hits = wbcast(w, n_hits, 0);
        wsync(w);
        filter_blocks_hits_copy(w, bs.out_el, bs.chain_off, nch, keep_chain, hits, hits2);
        H = hits2;
    }
    if (dbg_hits)
    {
        u32 k = (u32)n_hits < dbg_hits_cap ? (u32)n_hits : dbg_hits_cap;
        for (u32 i = (u32)w.lane; i < k; i += (u32)w.nl) dbg_hits[i] = H[i];
        if (w.lane == 0) *dbg_nhits = (u32)n_hits;
    }
    wsync(w);
    LNR_LAP(cnt, 6, tl);
    if (n_hits < 2) return 3;                // path_dst :1457
    // ---- _filterHits (:1417): drop hits whose own window distance >= reject (50). The reference compacts in place
    // and re-attaches the end flag of a dropped chain end to the hit kept before it (the sentinel if there is none):
    // here 32 hits at a time, kept hits to out[1 + number kept before], dropped ends OR their flag into out[number kept
    // before].
    if (w.lane == 0) { out[0] = H[0]; cnt.hits += (u64)(n_hits - 1); }
    wsync(w);
    int kept = 0;
    for (int c = 1; c < n_hits; c += w.nl)
    {
        const int it = c + w.lane;
        const bool valid = it < n_hits;
        u64 h = valid ? H[it] : 0;
        bool keep = false;
        if (valid)
        {
            u32 strand = (u32)cord_strand(h), id = (u32)cord_id(h);
            u64 x1 = cord_y(h) >> 4, x2 = cord_x(h) >> 4;
            u32 dist;   // _windowDist :676 (the bounds-checked wrapper; -f 1 checks the first index only, :693)
            if (in.ft == 1) dist = (x1 < in.nf1 && x2 < in.nf2[id]) ? window_dist32(in.s1[strand], in.nf1, x1, in.s2[id], in.nf2[id], x2) : 1000u;
            else dist = (x1 + 4 < in.nf1 && x2 + 4 < in.nf2[id]) ? window_dist48(in.f1[strand] + x1, in.f2[id] + x2) : 1000u;
            keep = dist < (u32)kWinReject;
        }
        const u32 bk = wballot(w, keep);
        const int before = kept + popc_below(w, bk);
        if (keep) out[1 + before] = h;
        wsync(w);
        if (valid && !keep && is_end(h)) wor_flag64(&out[before], kFlagEnd);
        kept += popc32(bk);
        wsync(w);
    }
    n_hits = 1 + kept;
    LNR_LAP(cnt, 7, tl);
    return 0;
}

LNR_PIPE int phase_map(const Warp & w, Arena & ar, u32 * hist256, u32 * bins, const PipeIn & in, u64 * A, u64 * B, int n,
                       u64 read_str, u64 read_end, int score_type, u64 * cords, int & n_cords, int cords_cap,
                       u64 * dbg_hits, u32 * dbg_nhits, u32 dbg_hits_cap, PipeCounters & cnt,
                       u64 * hits_out = (u64 *)0, u32 * n_hits_out = (u32 *)0, bool force_fit = true)
{
    // hits_out != null: stop after _filterHits and hand the hits over (it may alias A; capacity n) -- the window
    // extension then runs in its own thread-per-read kernel. hits_out == null: run path_dst_2 here.
    // Return 2 = nothing was touched because even the upper bound of the scratch need (all sizes <= n) does not
    // fit this arena: the caller re-runs the task with a larger arena (the anchors in A/B are consumed by a run).
    if (!force_fit && phase_map_scratch_bound(n) > ar.cap) return 2;
    arena_reset(ar);
    long long tl = LNR_CLOCK();
    if (n_hits_out && w.lane == 0) *n_hits_out = 0;
    if (dbg_nhits && w.lane == 0) *dbg_nhits = 1;
    if (dbg_hits && w.lane == 0 && dbg_hits_cap > 0) dbg_hits[0] = kFlagEnd;
    u64 * X; int n2;
    int rc = hits_sec_sort(w, ar, hist256, bins, A, B, n, cnt, tl, X, n2);
    if (rc) return rc == 3 ? 0 : rc;
    u64 * hits = arena_alloc<u64>(ar, (u64)n2 + 2);
    i32 * hits_score = arena_alloc<i32>(ar, (u64)n2 + 2);
    if (ar.failed) return 1;
    int n_hits = 1;
    rc = hits_sec_chain(w, ar, in, X, n2, score_type, hits, hits_score, n_hits, cnt, tl);
    if (rc) return rc;
    u64 * H = hits_out ? hits_out : arena_alloc<u64>(ar, (u64)n_hits + 2);
    if (ar.failed) return 1;
    rc = hits_sec_blocks(w, ar, hist256, in, hits, hits_score, n_hits, H, dbg_hits, dbg_nhits, dbg_hits_cap, cnt, tl);
    if (rc) return rc == 3 ? 0 : rc;
    if (hits_out)
    {
        if (w.lane == 0) *n_hits_out = (u32)n_hits;
        wsync(w);
        return 0;
    }
    int ok = path_dst_2(w, in, H, n_hits, cords, n_cords, cords_cap, read_str, read_end, cnt) ? 1 : 0;
    LNR_LAP(cnt, 8, tl);
    return ok ? 0 : 1;
}

// ----------------------------------------------------------------------------------------------------
// Phase 2 (lane 0): clean_blocks_, gather_blocks_, gather_gaps_y_ and the re-map decision
// (pmpfinder.cpp:2744-2749). gaps: forward-strand [y1, y2) intervals to re-map. Returns 1 if re-map is
// needed, 0 if not, -1 on capacity failure.
// ----------------------------------------------------------------------------------------------------
LNR_HD int phase_mid(u64 L, u64 * cords, int & n_cords, YPair * str_ends, Blk * sep, int & n_sep, YPair * gaps, int & n_gaps, int gaps_cap,
                     u32 win = kWin)
{
    i64 drop_len = imin64(2, (i64)((double)L * 0.05 / (double)win));
    n_cords = clean_blocks(cords, n_cords, (u64)drop_len, 50);
    int ns = 0;
    n_sep = gather_blocks(cords, n_cords, str_ends, ns, sep, 0, 1, (u32)n_cords, L, 1000, win, 1);
    if (n_cords < 2) ns = 0;
    int gap_sum = gather_gaps_y(str_ends, ns, gaps, n_gaps, gaps_cap, L, 1000);
    if (n_gaps >= gaps_cap) return -1;
    for (int i = 0; i < n_gaps; i++)   // getUPForwardy of every gap (:2753)
    {
        u64 a, b;
        forward_y(gaps[i], L, a, b);
        gaps[i].first = a; gaps[i].second = b;
    }
    return ((float)gap_sum / (float)L >= 0.7f) ? 1 : 0;
}

// ----------------------------------------------------------------------------------------------------
// Phase 3 (lane 0): chainApxCordsBlocks -> chainBlocksCords (:1747, cluster_util.cpp:1068), clean_blocks_,
// final flags (:2788-2801). sep/n_sep: blocks from the last gather_blocks_. tmp: cord scratch of cords_cap.
// ----------------------------------------------------------------------------------------------------
LNR_HD int phase_finish(u64 L, u64 * cords, int & n_cords, Blk * sep, int n_sep, Blk * sep2, i32 * score1, i32 * score2,
                        BlockScratch & s1, BlockScratch & s2, u64 * tmp, u32 win = kWin)
{
    for (int i = 0; i < n_sep; i++) sep2[i] = sep[i];
    int n1 = chain_blocks_single_strand(cords, sep, n_sep, s1, score1, 0, L, 16);
    int n2 = chain_blocks_single_strand(cords, sep2, n_sep, s2, score2, 1, L, 16);
    int best = best_strand_of(s1.out_el, s1.chain_off, n1, s2.out_el, s2.chain_off, n2);
    BlockScratch & sb = best == 0 ? s1 : s2;
    int nb = best == 0 ? n1 : n2;
    if (nb > 0)
    {
        revert_chain_block_strand(sb.out_el, sb.chain_off, nb, cords, best);
        int no = filter_blocks_cords(sb.out_el, sb.chain_off, nb, cords, tmp, 2);
        for (int i = 0; i < no; i++) cords[i] = tmp[i];
        n_cords = no;
    }
    i64 drop_len = imin64(2, (i64)((double)L * 0.05 / (double)win));
    n_cords = clean_blocks(cords, n_cords, (u64)drop_len, 50);
    int seg = 0;
    for (int i = 0; i < n_cords; i++)
    {
        u64 c = cords[i];
        if (seg) c |= kFlagRecd; else c &= ~kFlagRecd;
        c |= kFlagMain;
        if (is_end(c)) seg = 1 - seg;
        cords[i] = c;
    }
    return 0;
}

// warp versions of phase 2 / phase 3: the linear passes run on all lanes, the small block-chaining logic on lane 0
LNR_PIPE int phase_mid_w(const Warp & w, u64 L, u64 * cords, int & n_cords, YPair * str_ends, Blk * sep, int & n_sep, YPair * gaps,
                         int & n_gaps, int gaps_cap, u32 win = kWin)
{
    i64 drop_len = imin64(2, (i64)((double)L * 0.05 / (double)win));
    n_cords = clean_blocks_w(w, cords, n_cords, (u64)drop_len, 50);
    int ns = 0;
    n_sep = gather_blocks_w(w, cords, n_cords, str_ends, ns, sep, L, 1000, win, 1);
    int res = 0;
    if (w.lane == 0)
    {
        int gap_sum = gather_gaps_y(str_ends, ns, gaps, n_gaps, gaps_cap, L, 1000);
        if (n_gaps >= gaps_cap) res = -1;
        else
        {
            for (int i = 0; i < n_gaps; i++)
            {
                u64 a, b;
                forward_y(gaps[i], L, a, b);
                gaps[i].first = a; gaps[i].second = b;
            }
            res = ((float)gap_sum / (float)L >= 0.7f) ? 1 : 0;
        }
    }
    res = wbcast(w, res, 0);
    n_gaps = wbcast(w, n_gaps, 0);
    wsync(w);
    return res;
}

LNR_PIPE void phase_finish_w(const Warp & w, u64 L, u64 * cords, int & n_cords, Blk * sep, int n_sep, Blk * sep2, i32 * score1, i32 * score2,
                             BlockScratch & s1, BlockScratch & s2, u64 * tmp, u32 win = kWin)
{
    if (w.lane == 0)
    {
        for (int i = 0; i < n_sep; i++) sep2[i] = sep[i];
        int n1 = chain_blocks_single_strand(cords, sep, n_sep, s1, score1, 0, L, 16);
        int n2 = chain_blocks_single_strand(cords, sep2, n_sep, s2, score2, 1, L, 16);
        int best = best_strand_of(s1.out_el, s1.chain_off, n1, s2.out_el, s2.chain_off, n2);
        BlockScratch & sb = best == 0 ? s1 : s2;
        int nb = best == 0 ? n1 : n2;
        if (nb > 0)
        {
            revert_chain_block_strand(sb.out_el, sb.chain_off, nb, cords, best);
            int no = filter_blocks_cords(sb.out_el, sb.chain_off, nb, cords, tmp, 2);
            for (int i = 0; i < no; i++) cords[i] = tmp[i];
            n_cords = no;
        }
    }
    n_cords = wbcast(w, n_cords, 0);
    wsync(w);
    i64 drop_len = imin64(2, (i64)((double)L * 0.05 / (double)win));
    n_cords = clean_blocks_w(w, cords, n_cords, (u64)drop_len, 50);
    // main / record flags (pmpfinder.cpp:2788-2801): the record bit alternates after every block end
    int ends_before = 0;
    for (int c0 = 0; c0 < n_cords; c0 += w.nl)
    {
        int i = c0 + w.lane;
        u64 c = i < n_cords ? cords[i] : 0;
        u32 m = wballot(w, i < n_cords && is_end(c));
        int seg = (ends_before + popc_below(w, m)) & 1;
        if (i < n_cords)
        {
            if (seg) c |= kFlagRecd; else c &= ~kFlagRecd;
            cords[i] = c | kFlagMain;
        }
        ends_before += popc32(m);
    }
    wsync(w);
}

// ----------------------------------------------------------------------------------------------------
// -c 0 (apxMap with f_chain = 0, alg_type 1; pmpfinder.cpp:2773-2787): the anchors are grouped by sorting instead of the
// chaining DP. getDAnchorMatchList :2301 = getDAnchorList :2185 + getDHitList :2246, then path_dst_1 :1269.
// ----------------------------------------------------------------------------------------------------
// single IEEE operations, never contracted or approximated: the reference's float compares decide which anchors group
LNR_HD float c0_fmul(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
LNR_HD float c0_fdiv(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
LNR_HD u64 c0_scratch_bound(int n) { return ((u64)n + 4) * 24 + 1024; }

// getDAnchorList over the ascending anchors S[0..n) (S[0] = 0, the Anchors::init entry): a sequential recurrence -- the
// running medians ak2 / ak3 depend on every decision before -- that all lanes follow in step (the loads are broadcasts);
// lane 0 records the accepted runs as (c_b << 40) + (sb << 20) + k. The reference sorts an accepted run by y on the spot
// (sortPos2); nothing later in the scan reads below k, so here only the runs that become hits are sorted, by the caller.
// float thresholds exactly as the reference declares them (0.001f, 0.01f, 0.2f).
LNR_PIPE int c0_anchor_list(const Warp & w, const u64 * S, int n, u64 read_str, u64 read_end, u64 * list)
{
    const float dens = 0.001f, lens_rate = 0.01f, err = 0.2f;
    const i64 shape_len = 21;                                   // GlobalParms::shape_len base.cpp:99
    const i64 accept_lens = (i64)c0_fmul(lens_rate, (float)(read_end - read_str));
    int nl = 0;
    if (n <= 1) return 0;
    u64 ak2 = S[0], ak3 = S[0];
    i64 c_b = shape_len, sb = 1;
    u64 min_y = ~0ULL, max_y = 0;
    u64 prev = S[0];
    for (int k = 1; k < n; k++)
    {
        const u64 ak = S[k];
        const i64 anc_y = (i64)cord_y(ak);
        const i64 dy2 = iabs64(anc_y - (i64)cord_y(ak2)), dy3 = iabs64(anc_y - (i64)cord_y(ak3));
        const bool f_cont = (float)cord_x(ak - ak2) < c0_fmul(err, (float)dy2) || (float)cord_x(ak - ak3) < c0_fmul(err, (float)dy3);
        if (f_cont)
        {
            i64 dy = iabs64((i64)cord_y(ak) - (i64)cord_y(prev));
            c_b += imin64(dy, shape_len);
            ak2 = S[(sb + k) >> 1];
            ak3 = S[k - ((k - sb) >> 2)];
            const u64 y = cord_y(ak);
            min_y = y < min_y ? y : min_y;
            max_y = y > max_y ? y : max_y;
        }
        if (!f_cont || k == n - 1)
        {
            if (c_b > accept_lens && (i64)k - sb >= (i64)(u32)c0_fmul((float)(max_y - min_y), dens))
            {
                if (w.lane == 0) list[nl] = (u64)((c_b << 40) + (sb << 20) + (i64)k);
                nl++;
            }
            sb = k;
            ak2 = ak; ak3 = ak;
            c_b = shape_len;
            min_y = max_y = cord_y(ak);
        }
        prev = ak;
    }
    wsync(w);
    return nl;
}

// path_dst_1 :1269 with initCord :1178, nextCord :1209, endCord :1200 over the hits H[1..nh) (H[0] the sentinel), into an
// EMPTY cord list. Warp-uniform; the window extension is the one of path_dst_2. Returns false on overflow. max_len = the
// low 20 bits of cords[0] (Cord::setMaxLen cords.cpp:122), what apxMap compares with thd_sen.
template <int FT = 0>
LNR_PIPE bool path_dst_1(const Warp & w, const PipeIn & in, const u64 * H, int nh, u64 * cords, int & nc, int cap, u64 read_str,
                         u64 read_end, PipeCounters & cnt, u64 & max_len)
{
    const u64 L = in.L, cs = FT == 1 ? (u64)kWin32 : (FT == 2 ? (u64)kWin : (u64)in.win);
    max_len = 0;
    nc = 0;
    if (nh < 2) return true;                                    // path_dst :1457 isHitsEmpty
    if (cap < 2) return false;
    u64 c0 = kFlagEnd;                                          // cords[0]; rewritten by setMaxLen at the end
    int it = 1;
    u64 last = H[it++];
    if (w.lane == 0) { cords[0] = c0; cords[1] = last; }
    nc = 2;
    int pre = 1;
    while (true)
    {
        wsync(w);
        const u64 strand = cord_strand(last);
        u64 cordy_str = strand ? L - read_end : read_str;
        const u64 cordy_end = strand ? L - read_str - 1 : read_end;
        const u64 before = cords[nc - 2];
        const u64 pre_y = is_end(before) ? 0 : cord_y(before) + 1;
        cordy_str = pre_y > cordy_str ? pre_y : cordy_str;
        if (!extend_window<FT>(w, in, cords, nc, cap, last, cordy_str, cordy_end, cnt)) return false;
        // nextCord: the first later hit that starts a new block or lies behind the last cord, passes the window test and
        // stays inside [read_str, read_end)
        bool found = false, f_new_block = false;
        while (it < nh)
        {
            if (is_end(H[it - 1]))
            {
                last |= kFlagEnd;
                if (w.lane == 0) cords[nc - 1] = last;
                pre = nc;
                f_new_block = true;
            }
            const u64 h = H[it++];
            if (cord_y(h) > cord_y(last) || f_new_block)
            {
                const u32 st = (u32)cord_strand(h), id = (u32)cord_id(h);
                const u64 x1 = cord_y(h) >> 4, x2 = cord_x(h) >> 4;
                u32 dist;   // _windowDist :676, the bounds-checked wrapper (as in _filterHits)
                if ((FT ? FT : in.ft) == 1) dist = (x1 < in.nf1 && x2 < in.nf2[id]) ? window_dist32(in.s1[st], in.nf1, x1, in.s2[id], in.nf2[id], x2) : 1000u;
                else dist = (x1 + 4 < in.nf1 && x2 + 4 < in.nf2[id]) ? window_dist48(in.f1[st] + x1, in.f2[id] + x2) : 1000u;
                cnt.windows++;
                const u64 nyf = st ? L - 1 - cord_y(h) : cord_y(h);
                if (dist < (u32)kWinThr && cord_y(h) + cs < L && nyf >= read_str && nyf + cs < read_end)
                {
                    if (nc >= cap) return false;
                    last = h;
                    if (w.lane == 0) cords[nc] = h;
                    nc++;
                    found = true;
                    break;
                }
            }
        }
        if (!found)
        {
            if (f_new_block) { last |= kFlagEnd; pre = nc; }
            break;
        }
    }
    last |= kFlagEnd;                                           // set_cord_end(back(cords)), endCord
    const u64 len = (u64)(nc - pre);
    if (len > (c0 & kMaskY)) c0 = len + (c0 & ~kMaskY);
    if (w.lane == 0) { cords[nc - 1] = last; cords[0] = c0; }
    wsync(w);
    max_len = c0 & kMaskY;
    return true;
}

// One attempt of apxMap_ with alg_type 1 (:2632) on the raw anchors of a whole-read seeding task: A / B = the task's two
// n-entry regions (A[0] the sentinel slot). list_n / best_n = GetDHitListParms (thd_list_n, thd_best_n). Leaves the cords
// of the attempt in cords[0..nc). Return: 0 ok, 1 arena / cord capacity exhausted, 2 = nothing touched, arena too small.
template <int FT = 0>
LNR_PIPE int phase_c0(const Warp & w, Arena & ar, u32 * hist256, const PipeIn & in, u64 * A, u64 * B, int n, int list_n, int best_n,
                      u64 * cords, int & nc, int cap, PipeCounters & cnt, u64 & max_len, bool force_fit)
{
    nc = 0; max_len = 0;
    if (n >= (1 << 20)) return 1;                               // the run list packs anchor indices into 20 bits (:2232, :2264)
    if (!force_fit && c0_scratch_bound(n) > ar.cap) return 2;
    arena_reset(ar);
    if (w.lane == 0) A[0] = 0;                                  // Anchors::init(1) base.cpp:272
    wsync(w);
    if (n <= 1) return 0;
    const u64 read_str = 0, read_end = (u64)in.L & kMaskY;      // apxMap :2778: map_end = length(read), get_cord_y keeps 20 bits
    u64 * C = arena_alloc<u64>(ar, (u64)n);
    u64 * list = arena_alloc<u64>(ar, (u64)n);
    u64 * hits = arena_alloc<u64>(ar, (u64)n + 2);
    u64 * top = arena_alloc<u64>(ar, 32);
    if (ar.failed) return 1;
    u64 * S = radix_sort(w, hist256, A, B, C, n, 62, sort_key(0));          // anchors.sort = ska_sort, ascending
    u64 * T0 = (S == A) ? B : A;                                // the two buffers of {A, B, C} that are not S
    u64 * T1 = (S == C) ? B : C;
    if (T1 == T0) T1 = C;
    const int nl = c0_anchor_list(w, S, n, read_str, read_end, list);
    if (nl == 0) return 0;
    // getDHitList :2246: the runs by descending (c_b, sb, k); at most list_n are looked at, best_n become hits
    u64 * SL = radix_sort(w, hist256, list, T0, T1, nl, 64, sort_key(6));
    const int tmp = nl > list_n ? list_n : nl;                  // <= 20
    for (int i = w.lane; i < tmp; i += w.nl) top[i] = SL[i];
    wsync(w);
    int nh = 1, record_num = 1;
    if (w.lane == 0) hits[0] = kFlagEnd;
    const i64 l0 = (i64)top[0];
    for (int k = 0; k < tmp; k++)
    {
        if (record_num > best_n) break;
        const i64 lk = (i64)top[k];
        if (!((l0 / 10) < lk && lk)) break;
        const int sb = (int)((lk >> 20) & (i64)kMaskY), sc = (int)(lk & (i64)kMaskY);
        const int len = sc - sb;
        if (len > 0)
        {
            u64 * R = gnu_sort_w(w, hist256, S + sb, T0, T1, len, 20, sort_key(5));   // sortPos2: std::sort by y, its tie order
            for (int i = w.lane; i < len; i += w.nl) hits[nh + i] = hit2cord_dstr(R[i]);
            nh += len;
        }
        wsync(w);
        if (w.lane == 0) hits[nh - 1] |= kFlagEnd;              // back(hits): the sentinel when the run is empty
        wsync(w);
        ++record_num;
    }
    cnt.hits += (u64)(nh - 1);
    return path_dst_1<FT>(w, in, hits, nh, cords, nc, cap, read_str, read_end, cnt, max_len) ? 0 : 1;
}

// apxMap :2779-2786: the attempt is kept when its longest block reaches thd_sen (0.7) of the read, counted in windows
LNR_HD bool c0_attempt_too_short(u64 max_len, u64 L, u32 win)
{
    const float sen_thr = c0_fdiv(0.7f, (float)(i64)win);
    return (float)max_len < c0_fmul((float)L, sen_thr);
}

// clean_blocks_ with its default map error (:2787) and the main / record flags (:2788-2801)
LNR_PIPE void c0_finish(const Warp & w, u64 L, u64 * cords, int & nc, u32 win)
{
    i64 drop_len = imin64(2, (i64)((double)L * 0.05 / (double)win));
    nc = clean_blocks_w(w, cords, nc, (u64)drop_len, 50);
    int ends_before = 0;
    for (int c0 = 0; c0 < nc; c0 += w.nl)
    {
        int i = c0 + w.lane;
        u64 c = i < nc ? cords[i] : 0;
        u32 m = wballot(w, i < nc && is_end(c));
        int seg = (ends_before + popc_below(w, m)) & 1;
        if (i < nc)
        {
            if (seg) c |= kFlagRecd; else c &= ~kFlagRecd;
            cords[i] = c | kFlagMain;
        }
        ends_before += popc32(m);
    }
    wsync(w);
}

}  // namespace lnr
