// HIndex (-i 2): createHIndex (index_util.cpp:1471) rebuilt as sample extraction -> device-wide LSD radix sort ->
// block assembly -> open-addressing directory; and getHIndexMatchAll (pmpfinder.cpp:1918) as count / scan / fill kernels.
// Included by lnr_kernels.cu (uses its context, scan helper and accessors). Net semantics follow SURVEY App. F, which
// the CPU oracle (oracle/lnr_oracle_apx.inc build_hindex / seed_hindex) pins against the reference at -t 1/4/8:
//   * per contig T chunks (index_util.cpp:742-760), samples at k % 8 == 0, a (head, body) pair whenever X changes;
//     the last pair of a chunk is filed under the X of the chunk's final k-mer (:793)
//   * pairs grouped by X ascending, bodies descending; head = (bodies+1) << 40 | X; blocks < 1024 lose Y bits 41..60
//   * directory: X -> first body for small blocks; virtual head + one entry per run of equal Y for blocks >= 1024
// ACGT genomes only (the reference's N branch, :767-775, is documented as broken there and not restated).
#pragma once
#include "lnr_radix.cuh"

static const int kSpanH = 17, kStepH = 8;
static const u64 kBlockLimitH = 1024;

struct HChunk
{
    u64 base_off; i64 len;
    i64 start, csize;        // k in [start, start + csize)
    i64 k_first;             // first k with k % 8 == 0
    i64 n_samples;
    u64 sample0;
    u32 contig;
    u32 x_last;              // X of the chunk's final k-mer (k = start + csize - 1)
    i64 m_last_emit;         // index of the last emitted sample of the chunk (-1: none)
};

__device__ __forceinline__ void hidx_eval(const GAcc & acc, const HChunk & ch, i64 k, SeedVal & sv)
{
    // hashInit at `start`, hashNext from k = start: consistent windows from the first step (no N)
    eval_sample_t<kSpanH, true>(acc, k, kSpanH, ch.start, ch.start + kSpanH - 1, 0, sv);
}

__global__ void k_hidx_prep(const u8 * __restrict__ g, HChunk * chunks, u32 n_chunks)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chunks) return;
    HChunk ch = chunks[i];
    GAcc acc = {g + ch.base_off, ch.len};
    SeedVal sv;
    hidx_eval(acc, ch, ch.start + ch.csize - 1, sv);
    ch.x_last = sv.X;
    // last emitted sample: last m whose X differs from its predecessor's (sample 0 always emits)
    i64 m = ch.n_samples - 1;
    if (m >= 0)
    {
        hidx_eval(acc, ch, ch.k_first + kStepH * m, sv);
        u32 X = sv.X;
        while (m > 0)
        {
            hidx_eval(acc, ch, ch.k_first + kStepH * (m - 1), sv);
            if (sv.X != X) break;
            m--;
        }
    }
    ch.m_last_emit = m;
    chunks[i] = ch;
}

__global__ void k_check_acgt(const u8 * __restrict__ g, const u64 * __restrict__ off, const u64 * __restrict__ len, u32 n_contigs, u32 * bad)
{
    for (u32 c = 0; c < n_contigs; c++)
        for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < len[c]; i += (u64)gridDim.x * blockDim.x)
            if (g[off[c] + i] > 3) { atomicAdd(bad, 1u); return; }
}

// one thread per sample; emitted pairs are appended (any order: the sort fixes it). Two forms share the evaluation:
// xhist != null: count the emitted pairs per X (X < 2^18 for the 9-base minimizer of a 17-base shape), nothing is stored --
// the sharded build cuts the X axis into ranges of equal pair count from it; xhist == null: store the pairs whose X lies in
// [x_lo, x_hi) (the whole axis for the single-GPU build).
static const u32 kXRangeH = 1u << 18;
__global__ void __launch_bounds__(256) k_hidx_pairs(const u8 * __restrict__ g, const HChunk * __restrict__ chunks, u32 n_chunks, u64 n_samples,
                                                    u64 * __restrict__ body, u32 * __restrict__ xkey, unsigned long long * n_pairs,
                                                    u32 x_lo, u32 x_hi, u32 * __restrict__ xhist)
{
    u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool emit = false;
    u64 b = 0; u32 X = 0;
    if (s < n_samples)
    {
        u32 lo = 0, hi = n_chunks;
        while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (chunks[mid].sample0 <= s) lo = mid; else hi = mid; }
        const HChunk ch = chunks[lo];
        GAcc acc = {g + ch.base_off, ch.len};
        i64 m = (i64)(s - ch.sample0);
        i64 k = ch.k_first + kStepH * m;
        SeedVal sv;
        hidx_eval(acc, ch, k, sv);
        emit = true;
        if (m > 0) { SeedVal pv; hidx_eval(acc, ch, k - kStepH, pv); emit = pv.X != sv.X; }
        if (emit)
        {
            X = (m == ch.m_last_emit) ? ch.x_last : sv.X;
            b = (1ULL << 63) | (((u64)sv.Y << 41) + ((u64)ch.contig << 30) + (u64)k);
            if (sv.strand) b |= 1ULL << 40;
            if (xhist) { atomicAdd(&xhist[X < kXRangeH ? X : kXRangeH - 1], 1u); emit = false; }
            else emit = X >= x_lo && X < x_hi;
        }
    }
    if (xhist) return;
    u32 m = __ballot_sync(0xffffffffu, emit);
    u64 base = 0;
    if ((threadIdx.x & 31) == 0 && m) base = atomicAdd(n_pairs, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (emit)
    {
        u64 o = base + __popc(m & ((1u << (threadIdx.x & 31)) - 1));
        body[o] = b;
        xkey[o] = X;
    }
}

// after the sort (X ascending, body descending): block starts
__global__ void __launch_bounds__(256) k_hidx_flags(const u32 * __restrict__ xkey, u64 n, u32 * __restrict__ flag)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) flag[i] = (i < n && (i == 0 || xkey[i] != xkey[i - 1])) ? 1u : 0u;
}
// bid[i] = exclusive scan of flag = number of block starts before i  (block number of element i = bid[i] + flag[i] - 1)
__global__ void __launch_bounds__(256) k_hidx_starts(const u32 * __restrict__ flag, const u64 * __restrict__ bid, u64 n, u64 * __restrict__ starts)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) starts[bid[i]] = i;
    if (i == n) starts[bid[n]] = n;   // sentinel: end of the last block
}
// ysa[i + block(i) + 1] = body, head in front of every block; small blocks lose Y bits 41..60 (:1414-1417)
__global__ void __launch_bounds__(256) k_hidx_write(const u64 * __restrict__ body, const u32 * __restrict__ xkey, const u32 * __restrict__ flag,
                                                    const u64 * __restrict__ bid, const u64 * __restrict__ starts, u64 n, u64 * __restrict__ ysa)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 blk = bid[i] + flag[i] - 1;
    u64 ptr = starts[blk + 1] - starts[blk] + 1;
    u64 v = body[i];
    if (ptr < kBlockLimitH) v &= ~(0xfffffULL << 41);
    ysa[i + blk + 1] = v;
    if (flag[i]) ysa[i + blk] = (ptr << 40) + (u64)xkey[i];
}

// ---- directory: open addressing on 16-byte nodes {val1 = key << 2 | type, val2}; probing as the reference (:1006-1034)
struct HNode { unsigned long long val1; u32 val2; u32 pad; };
__device__ __forceinline__ u64 xnode_hash(u64 key)   // XNodeFunc::hash index_util.cpp:971
{
    key = (~key) + (key << 21);
    key = key ^ (key >> 24);
    key = (key + (key << 3)) + (key << 8);
    key = key ^ (key >> 14);
    key = (key + (key << 2)) + (key << 4);
    key = key ^ (key >> 28);
    key = key + (key << 31);
    return key;
}
__device__ __forceinline__ void hdir_insert(HNode * tab, u64 mask, u64 key, u32 val2, u64 type)
{
    u64 h = xnode_hash(key) & mask, delta = 0;
    unsigned long long v1 = (key << 2) + type;
    while (atomicCAS(&tab[h].val1, 0ULL, v1) != 0ULL) { h = (h + delta + 1) & mask; delta++; }
    tab[h].val2 = val2;
}
// getXDir (index_util.cpp:1102): exact-match lookups; returns empty_dir on a miss
__device__ __forceinline__ u64 hdir_lookup(const HNode * __restrict__ tab, u64 mask, u64 X, u64 Y, u64 empty_dir)
{
    // A virtual head (block >= 1024) sends the lookup on to the key (Y, X). For Y == 0 that is the key it started with: the
    // reference then probes the same slots again, for ever (index_util.cpp:1082-1086 -- `linear -i 2` does not return from
    // such a read; from ~200 Mbase on every run meets one). The (key -> value) map has no such entry, so the canonical
    // answer is a miss, which is what the oracle's exact-match map returns.
    u64 val = (X << 2) + 1, delta = 0;
    u64 h = xnode_hash(X) & mask;
    bool rehashed = false;
    while (true)
    {
        u64 v1 = tab[h].val1;
        if (!v1) return empty_dir;
        u64 c = v1 ^ val;
        if (c == 0) return tab[h].val2;
        if (c == 2)
        {
            if (rehashed || Y == 0) return empty_dir;
            rehashed = true;
            val = (Y << 42) + (X << 2) + 1; h = xnode_hash((Y << 40) + X) & mask; delta = 0; continue;
        }
        h = (h + delta + 1) & mask; delta++;
    }
}
// entries per block: 1, or 1 + number of Y runs for blocks >= 1024. mode 0 = count, 1 = insert
__global__ void __launch_bounds__(256) k_hidx_dir(const u64 * __restrict__ ysa, const u64 * __restrict__ starts, u64 n_blocks, int mode,
                                                  unsigned long long * n_entries, HNode * tab, u64 mask)
{
    u64 blk = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 cnt = 0;
    if (blk < n_blocks)
    {
        u64 i = starts[blk] + blk;             // head position in ysa
        u64 ptr = starts[blk + 1] - starts[blk] + 1;
        u64 X = ysa[i] & ((1ULL << 40) - 1);
        if (ptr < kBlockLimitH)
        {
            cnt = 1;
            if (mode) hdir_insert(tab, mask, X, (u32)(i + 1), 1);
        }
        else
        {
            cnt = 1;
            if (mode) hdir_insert(tab, mask, X, ~1u, 3);
            for (u64 j = i + 1; j < i + ptr; j++)
                if (((ysa[j] ^ ysa[j - 1]) >> 41) & 0xfffff)
                {
                    cnt++;
                    if (mode) hdir_insert(tab, mask, X + ((ysa[j] & ((1ULL << 61) - (1ULL << 41))) >> 1), (u32)j, 1);
                }
        }
    }
    if (!mode)
    {
        for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(n_entries, (unsigned long long)cnt);
    }
}
// The same directory entries derived from the assembled ysa alone (sharded build: the block starts of the other ranks'
// slices are not on this rank): one thread per ysa word; a head word (bit 63 clear -- every body carries bit 63) gives its
// block's X and length. n_words = pairs + blocks (the two terminators behind them are not looked at).
__global__ void __launch_bounds__(256) k_hidx_dir_scan(const u64 * __restrict__ ysa, u64 n_words, int mode, unsigned long long * n_entries,
                                                       HNode * tab, u64 mask)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 cnt = 0;
    if (i < n_words)
    {
        const u64 head = ysa[i];
        if (!(head >> 63))
        {
            const u64 ptr = head >> 40, X = head & ((1ULL << 40) - 1);
            cnt = 1;
            if (ptr < kBlockLimitH)
            {
                if (mode) hdir_insert(tab, mask, X, (u32)(i + 1), 1);
            }
            else
            {
                if (mode) hdir_insert(tab, mask, X, ~1u, 3);
                for (u64 j = i + 1; j < i + ptr; j++)
                    if (((ysa[j] ^ ysa[j - 1]) >> 41) & 0xfffff)
                    {
                        cnt++;
                        if (mode) hdir_insert(tab, mask, X + ((ysa[j] & ((1ULL << 61) - (1ULL << 41))) >> 1), (u32)j, 1);
                    }
            }
        }
    }
    if (!mode)
    {
        for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(n_entries, (unsigned long long)cnt);
    }
}
__global__ void k_hidx_dir_export(const HNode * __restrict__ tab, u64 len, u64 * __restrict__ kv, unsigned long long * n)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len || !tab[i].val1) return;
    u64 o = atomicAdd(n, 1ULL);
    kv[2 * o] = tab[i].val1;
    kv[2 * o + 1] = tab[i].val2;
}

// ---- seeding: getHIndexMatchAll (pmpfinder.cpp:1918). Samples k = str + alpha*m - 1 (m >= 1) while k < end - 17.
struct HIndexDev { const u64 * ysa; u64 n_ysa; u64 empty_dir; const HNode * tab; u64 mask; };

__device__ __forceinline__ void hseed_eval(const GAcc & acc, const SeedTask & t, u32 m, SeedVal & sv, u32 & k)
{
    i64 k0 = (i64)t.str;
    i64 kk = k0 + (i64)t.alpha * m - 1;
    eval_sample_t<kSpanH, false>(acc, kk, kk - k0 + 1, (i64)t.kskip, k0 + kSpanH - 1, t.bias, sv);
    k = (u32)kk;
}
__global__ void k_hseed_prep(const u8 * __restrict__ bases, const u64 * __restrict__ read_off, SeedTask * tasks, u32 n_tasks)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tasks) return;
    SeedTask t = tasks[i];
    u64 L = read_off[t.read + 1] - read_off[t.read];
    GAcc acc = {bases + read_off[t.read], (i64)L};
    t.kskip = (u32)hash_init_skip<kSpanH>(acc, 0, (i64)L);
    t.bias = selector_bias<kSpanH>(acc, (i64)t.kskip, (i64)t.str);
    tasks[i] = t;
}
// walks the raw ysa words from pos while their Y field is the query Y or 0 (head words included);
// FILL = false: count only
template <bool FILL>
__device__ __forceinline__ u32 hseed_scan(const HIndexDev & hx, u64 pos, u32 Y, u32 strand, u64 k, u64 L, u64 * out)
{
    u32 c = 0;
    const u64 idx_end = (1ULL << 40) - 1;     // getCordX(create_cord(MAX_ID, MAX_X, ..)) ; idx_str = 0
    while (true)
    {
        u64 w = __ldg(hx.ysa + pos);
        u32 wy = (u32)(w >> 41) & 0xfffff;
        if (wy != Y && wy != 0) break;
        u64 idx = w & ((1ULL << 40) - 1);
        if (idx < idx_end)
        {
            if (FILL)
            {
                u64 id = (idx >> 30) & 1023, x = idx & ((1ULL << 30) - 1);
                bool rev = (((w >> 40) & 1) ^ strand) != 0;
                u64 y = rev ? L - 1 - k : k;
                out[c] = create_cord(id, x - y + kAnchorZero, y, rev ? 1 : 0);   // make_anchor cords.cpp:319
            }
            c++;
        }
        if (++pos > hx.n_ysa - 1) break;
    }
    return c;
}
__global__ void __launch_bounds__(256) k_hseed_count(const u8 * __restrict__ bases, const u64 * __restrict__ read_off,
                                                     const SeedTask * __restrict__ tasks, u32 n_tasks, u64 n_samples, HIndexDev hx,
                                                     u64 * __restrict__ info, u32 * __restrict__ count, unsigned long long * counters)
{
    u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 c = 0;
    if (s < n_samples)
    {
        u32 ti = find_task(tasks, n_tasks, s);
        SeedTask t = tasks[ti];
        u32 m = (u32)(s - t.sample0) + 1;
        u64 L = read_off[t.read + 1] - read_off[t.read];
        GAcc acc = {bases + read_off[t.read], (i64)L};
        SeedVal sv; u32 k;
        hseed_eval(acc, t, m, sv, k);
        u32 xprev = 0;
        if (m > 1) { SeedVal pv; u32 kp; hseed_eval(acc, t, m - 1, pv, kp); xprev = pv.X; }
        u64 inf = 0;
        if (sv.X != xprev)
        {
            u64 pos = hdir_lookup(hx.tab, hx.mask, sv.X, sv.Y, hx.empty_dir);
            u64 ptr = (__ldg(hx.ysa + pos - 1) >> 40) & ((1ULL << 23) - 1);
            if (pos != hx.empty_dir && ptr < 64)
            {
                c = hseed_scan<false>(hx, pos, sv.Y, sv.strand, k, L, (u64 *)0);
                inf = pos | ((u64)sv.Y << 40) | ((u64)sv.strand << 48) | (1ULL << 63);
            }
        }
        info[s] = inf;
        count[s] = c;
    }
    u32 tot = c;
    for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if ((threadIdx.x & 31) == 0 && tot) { atomicAdd(&counters[1], (unsigned long long)tot); atomicAdd(&counters[2], (unsigned long long)tot); }
}
__global__ void __launch_bounds__(256) k_hseed_fill(const u64 * __restrict__ read_off, const SeedTask * __restrict__ tasks, u32 n_tasks,
                                                    u64 n_samples, HIndexDev hx, const u64 * __restrict__ info, const u64 * __restrict__ aoff,
                                                    u64 * __restrict__ anchors)
{
    u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_samples) return;
    u64 inf = info[s];
    if (!(inf >> 63)) return;
    u32 ti = find_task(tasks, n_tasks, s);
    SeedTask t = tasks[ti];
    u32 m = (u32)(s - t.sample0) + 1;
    u64 k = (u64)t.str + (u64)t.alpha * m - 1;
    u64 L = read_off[t.read + 1] - read_off[t.read];
    hseed_scan<true>(hx, inf & ((1ULL << 40) - 1), (u32)(inf >> 40) & 0xff, (u32)(inf >> 48) & 1, k, L, anchors + aoff[s] + ti + 1);
}
