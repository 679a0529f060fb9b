// sm_100a kernels + C ABI (include/lnr_b200.h) of the approximate-map path of `linear filter`.
//
// Data layout in HBM
//   genome    1 byte/base as delivered (Dna5 ordinals), contigs back to back at 256-byte aligned offsets with
//             >= 256 zero bytes after each contig
//   DIndex    dir: int32[2^26+1] bucket offsets; hs: uint64[n_hs] records, ascending inside a bucket
//   features  F96 (3 x int32) per 16 bases; per contig for the genome, per (read, strand) for a batch
//   batch     read bases back to back, anchors / scratch per read, cords per read at host-computed offsets
//
//   lookup    hsy: the Y byte of every hs record; dirx: per bucket 64 B = one DRAM burst (start, size, first 56 Y keys)
//
// Kernels (all HBM / issue / latency bound integer work, no tensor-core shaped math on this path)
//   k_feat_genome / k_feat_reads      2-mer/48 features: one thread per 16-base cell, 3-cell sums through smem; reads:
//                                     persistent CTAs, both strands of a tile per iteration (k_feat_pairs descriptors), a
//                                     cell = two aligned 16-byte loads, 2-bit packing, 4 lookups in a 5-base window table
//   k_idx_prep, k_idx_emit            DIndex samples: genome tile by one TMA bulk copy, packed window evaluation, emit rule
//                                     by block max-scan, dense (X, record) pairs
//   k_idx_partcount / k_idx_part / k_idx_count / k_idx_place / k_idx_rank   partition by X range, count / place / rank in L2
//   k_scan_*                          device-wide exclusive scan (reduce / spine / apply), fused bucket omission
//   k_idx_split_y, k_idx_dirx         Y byte array, 64-byte lookup entries
//   k_seed_prep / k_seed_count / k_seed_fill / k_seed_stats   per-read seeding, one thread per sample, exact-size output;
//                                     random index reads with a 64-byte L2 fetch hint (a plain miss brings in 128 bytes)
//   k_hits_sort / k_hits_chain / k_hits_blocks      the hit stage, one kernel per section (lnr_pipeline.h hits_sec_*):
//                                     persistent warp per read + atomic queue, heaviest tasks first
//   k_map_hits                        the same three sections in one kernel (re-map and big-arena passes)
//   k_map_extend                      window extension, one warp per read on a regular grid
//   k_map_finish                      warp per read: clean / gaps / re-map decision / cord-block chaining
//   k_gather_cords                    per-read cords -> caller's concatenated layout
//   lnr_ingest.cuh                    FASTA / FASTQ text -> ordinals, offsets, id spans (16 bytes per thread)
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/lnr_b200.h"
#include "lnr_core.h"
#include "lnr_pipeline.h"
#include "lnr_bamrec.h"

using namespace lnr;

// =====================================================================================================
// host-side context
// =====================================================================================================
// 8 CTAs per SM (64 registers): both kernels are latency bound, the two extra CTAs buy more than the spills cost
// (chain 1.78 -> 1.57 ms, blocks 1.38 -> 1.33 ms per 32 768 reads; 10 CTAs: no further gain)
#ifndef LNR_CHAIN_MIN_CTAS
#define LNR_CHAIN_MIN_CTAS 8
#endif
#ifndef LNR_BLOCKS_MIN_CTAS
#define LNR_BLOCKS_MIN_CTAS 8
#endif
struct KernelStat { double ms; uint64_t launches; };

struct DevBuf
{
    void * p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T * as() const { return (T *)p; }
};

// Pinned host staging for the small per-batch tables (offsets, task descriptors). They reach the device through a copy
// KERNEL that reads the pinned pages over PCIe, not through the copy engine: a small cudaMemcpyAsync queues behind any bulk
// upload another host thread has in flight on the same engine and would stall this thread's batch for the whole upload.
struct HostStage
{
    u8 * p = nullptr; size_t cap = 0, used = 0;
    cudaError_t reserve(size_t bytes)
    {
        used = 0;
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaHostAlloc((void **)&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void * take(size_t bytes)
    {
        size_t b = (bytes + 15) & ~(size_t)15;
        if (used + b > cap) return nullptr;
        void * r = p + used; used += b; return r;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; used = 0; }
};

struct lnr_ctx
{
    int device = 0;
    HostStage stage;
    int n_sm = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    bool profiling = false;
    uint64_t launches_total = 0;
    std::map<std::string, KernelStat> stats;
    std::vector<std::string> stat_order;
    struct Pending { std::string name; cudaEvent_t a, b; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> event_pool;
    // batch workspace
    DevBuf bases, read_off, tasks, sample_info, sample_cnt, scan_tmp, anchorsA, anchorsB, feats, foff, ftile,
        cords, cords_base, ncords, slots, bins, arena, tasks2, misc, out_cords, out_off, dbg_hits, dbg_hoff, dbg_nhits,
        dbg_c1, dbg_nc1, read_meta;
    uint64_t counters[8] = {0};
    uint64_t diag[8] = {0};
    uint64_t anchors_host_total = 0;   // raw anchors of the DIndex seeding passes of the current batch
    uint64_t stage_cycles[16] = {0};
    uint64_t longest_cycles[16] = {0};
    const void * bins_zeroed = nullptr;
    size_t bins_zeroed_cap = 0;
    DevBuf ing[8];   // read-ingest temporaries (lnr_ingest.cuh)
    void * reads_cache = nullptr; size_t reads_cache_bytes = 0;   // last output block given back by lnr_reads_destroy
    DevBuf packed;   // 2-bit packed batch as uploaded (lnr_apxmap_batch_packed)
    DevBuf remap_list, order, order2, mask_ctr, tile_read, task_nhits, task_state, big_arena, big_list, heavy_list, seed_masks, seed_mask_off, warp_rec, task_info, feat_pairs;
    size_t big_arena_bytes_per_warp = 128u << 20;
    int map_warps_per_cta = 4;
    int map_ctas_per_sm = 6;
    int sort_ctas_per_sm = 8, chain_ctas_per_sm = LNR_CHAIN_MIN_CTAS, blocks_ctas_per_sm = LNR_BLOCKS_MIN_CTAS;   // residency of the three hit-section kernels (their launch bounds)
    int extend_group = 32;
    size_t arena_bytes_per_warp = 4u << 20;
};

struct lnr_genome
{
    lnr_ctx * ctx;
    uint32_t n_contigs;
    std::vector<uint64_t> len, off;
    uint64_t total_padded;
    u8 * d_bases;   // owned
};
struct lnr_feats
{
    lnr_ctx * ctx;
    int feature_type;
    uint32_t n_contigs;
    std::vector<uint32_t> n;        // entries per contig
    std::vector<uint64_t> off;      // entry offset per contig
    F96 * d_f;                      // owned, all contigs back to back (feature_type 1: the same buffer holds shorts)
    const F96 ** d_ptrs;            // device table of per-contig pointers (feature_type 1: const i16 * behind the cast)
    u32 * d_n;                      // device table of counts
};
struct HNode;
struct lnr_index
{
    lnr_ctx * ctx;
    int index_type;          // 1 = DIndex, 2 = HIndex
    i32 * d_dir;
    u64 * d_hs;
    uint64_t n_hs;
    u8 * d_hsy = nullptr;    // low byte (the 8-bit Y key) of every hs record, split out for the seeding count pass
    uint4 * d_dirx = nullptr; // per bucket 64 B: start, size and the first 56 Y keys -- one DRAM access per seed for most buckets
    // HIndex (include/index_util.h:139-248)
    u64 * d_ysa = nullptr;
    uint64_t n_ysa = 0, empty_dir = 0;
    HNode * d_tab = nullptr;
    uint64_t tab_len = 0, n_dir_entries = 0;
};

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            char b_[512];                                                                               \
            snprintf(b_, sizeof b_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            ctx->err = b_;                                                                              \
            return LNR_E_CUDA;                                                                          \
        }                                                                                               \
    } while (0)

static int fail(lnr_ctx * ctx, int code, const char * msg) { if (ctx) ctx->err = msg; return code; }

// kernel launch with optional event bracketing (bench: per-kernel device time on the launching stream)
struct LaunchScope
{
    lnr_ctx * ctx; bool on; lnr_ctx::Pending p;
    LaunchScope(lnr_ctx * c, const char * name, int n_launches = 1) : ctx(c), on(c->profiling)
    {
        c->launches_total += (uint64_t)n_launches;
        if (!on) return;
        p.name = name;
        for (cudaEvent_t * e : {&p.a, &p.b})
        {
            if (!ctx->event_pool.empty()) { *e = ctx->event_pool.back(); ctx->event_pool.pop_back(); }
            else cudaEventCreate(e);
        }
        cudaEventRecord(p.a, ctx->stream);
    }
    ~LaunchScope()
    {
        if (!on) return;
        cudaEventRecord(p.b, ctx->stream);
        ctx->pending.push_back(p);
        if (getenv("LNR_TRACE_SYNC"))      // debugging aid: name every launch as it completes (the last line names the one before a hang)
        {
            cudaError_t e = cudaStreamSynchronize(ctx->stream);
            fprintf(stderr, "[lnr sync] %s %s\n", p.name.c_str(), e == cudaSuccess ? "done" : cudaGetErrorString(e));
            fflush(stderr);
        }
    }
};
static void harvest_stats(lnr_ctx * ctx)
{
    for (auto & p : ctx->pending)
    {
        cudaEventSynchronize(p.b);
        float ms = 0;
        cudaEventElapsedTime(&ms, p.a, p.b);
        auto it = ctx->stats.find(p.name);
        if (it == ctx->stats.end()) { ctx->stats[p.name] = KernelStat{0, 0}; ctx->stat_order.push_back(p.name); it = ctx->stats.find(p.name); }
        it->second.ms += ms;
        it->second.launches += 1;
        ctx->event_pool.push_back(p.a);
        ctx->event_pool.push_back(p.b);
    }
    ctx->pending.clear();
}

// =====================================================================================================
// device helpers
// =====================================================================================================
__device__ __forceinline__ u64 globaltimer_ns() { u64 t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
struct GAcc   // bounds-checked byte view of one sequence in global memory; out of range reads as 0
{
    const u8 * s; i64 len;
    __device__ __forceinline__ int operator()(i64 p) const { return (p >= 0 && p < len) ? (int)__ldg(s + p) : 0; }
};
struct GRcAcc   // reverse-complement view (_compltRvseStr base.cpp:335)
{
    const u8 * s; i64 len;
    __device__ __forceinline__ int operator()(i64 p) const
    {
        if (p < 0 || p >= len) return 0;
        int c = (int)__ldg(s + (len - 1 - p));
        return c < 4 ? 3 - c : 4;
    }
};

// =====================================================================================================
// features
// =====================================================================================================
static const u32 kMaskPools = 256;    // pools of the seeding match-mask allocator (k_seed_count)
static const int FT = 256;           // threads per CTA = cells per CTA
static const int FE = FT - 2;        // entries per CTA

// ---- table-driven cell for N-free interior cells ----------------------------------------------------------------
// The kernel is bound by the L1 data pipe (ncu: l1tex__data_pipe_lsu_wavefronts 84 % with 4-byte loads and a 4-base table),
// so a cell costs as few wavefronts as possible: its 17 bases arrive as the TWO aligned 16-byte blocks that hold them
// (feat_cell_issue; [q & ~15, +32) must be readable), are packed to 2 bits each (feat_cell_finish; Q: base i at bits
// 2(16-i); a cell with an N takes the general evaluation), and a 1024-entry shared table holds the field increments of
// the four 2-mers inside every 5-base window: a cell is 4 table adds. The reverse strand's cell is the pair-reversed
// complement of the forward 17 bases.
struct FeatTab { u64 lo[1024]; u32 hi[1024]; };
__device__ __forceinline__ void feat_add(u32 id, u64 & lo, u32 & hi)
{
    if (id < 10) lo += 1ULL << (6 * id);
    else if (id < 15) hi += 1u << (6 * (id - 10));
}
__device__ __forceinline__ void feat_tab_init(FeatTab & T)   // blockDim.x == 256, caller syncs
{
    for (u32 t = threadIdx.x; t < 1024; t += FT)
    {
        u64 lo = 0; u32 hi = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) feat_add((t >> (2 * (3 - k))) & 15u, lo, hi);   // bases k, k+1 of the window (base 0 on top)
        T.lo[t] = lo; T.hi[t] = hi;
    }
}
__device__ __forceinline__ u64 rc34(u64 Q)
{
    u64 y = __brevll(~Q) >> 30;
    return (((y >> 1) & 0x5555555555555555ULL) | ((y & 0x5555555555555555ULL) << 1)) & ((1ULL << 34) - 1);
}
__device__ __forceinline__ void feat_cell_tab(const FeatTab & T, u64 Q, u64 & lo, u32 & hi)
{
    lo = 0; hi = 0;
#pragma unroll
    for (int g = 0; g < 4; g++)
    {
        const u32 idx = (u32)(Q >> (24 - 8 * g)) & 0x3ffu;      // bases 4g .. 4g+4
        lo += T.lo[idx]; hi += T.hi[idx];
    }
}
// A cell in two steps, so that a thread can have the loads of several cells in flight before it needs the first word
// (k_feat_reads: the forward and the reverse cell of the same tile index): issue = decide the path and start the two
// 16-byte loads of the packed path; finish = pack, table adds -- or the general base-by-base evaluation (read ends, N).
struct CellLoad { uint4 v0, v1; unsigned a; int mode; };   // a = byte offset of the cell in the 32 bytes; mode 0: no cell, 1: loaded, 2: general evaluation
template <bool RC>
__device__ __forceinline__ void feat_cell_issue(const u8 * buf0, const u8 * s, i64 L, u32 c, u32 n_entries, CellLoad & cl)
{
    cl.mode = 0; cl.a = 0;
    cl.v0 = make_uint4(0, 0, 0, 0); cl.v1 = make_uint4(0, 0, 0, 0);
    if (c >= n_entries + 2) return;
    const i64 f0 = RC ? L - 17 - 16 * (i64)c : 16 * (i64)c;   // forward position of the cell's lowest base
    const u8 * q = s + f0;
    const uint4 * pv = (const uint4 *)((uintptr_t)q & ~(uintptr_t)15);
    if (f0 >= 0 && f0 + 32 <= L && (const u8 *)pv >= buf0)    // the 32 bytes lie inside this read (and inside the caller's buffer)
    {
        cl.mode = 1;
        cl.a = (unsigned)(uintptr_t)q & 15u;
        cl.v0 = __ldg(pv); cl.v1 = __ldg(pv + 1);
    }
    else cl.mode = 2;
}
template <bool RC>
__device__ __forceinline__ void feat_cell_finish(const CellLoad & cl, const u8 * s, i64 L, u32 c, const FeatTab & T, u64 & lo, u32 & hi)
{
    lo = 0; hi = 0;
    if (cl.mode == 0) return;
    if (cl.mode == 1)
    {
        // the 5 words that hold bytes a .. a+16: skip a >> 2 whole words (two select levels), then funnel by the byte rest
        const u32 W[8] = {cl.v0.x, cl.v0.y, cl.v0.z, cl.v0.w, cl.v1.x, cl.v1.y, cl.v1.z, cl.v1.w};
        const bool s2 = (cl.a & 8u) != 0, s1 = (cl.a & 4u) != 0;
        u32 V[6], U[5];
#pragma unroll
        for (int i = 0; i < 6; i++) V[i] = s2 ? W[i + 2] : W[i];
#pragma unroll
        for (int i = 0; i < 5; i++) U[i] = s1 ? V[i + 1] : V[i];
        const unsigned sh = (cl.a & 3u) * 8u;
        u32 A[4];
#pragma unroll
        for (int i = 0; i < 4; i++) A[i] = __funnelshift_r(U[i], U[i + 1], sh);
        const u32 b16 = (U[4] >> sh) & 0xffu;
        if (!((A[0] | A[1] | A[2] | A[3] | b16) & 0xfcfcfcfcu))
        {
            u32 hi32 = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) hi32 = (hi32 << 8) | ((A[i] * 0x40100401u) >> 24);
            const u64 Q = ((u64)hi32 << 2) | b16;
            feat_cell_tab(T, RC ? rc34(Q) : Q, lo, hi);
            return;
        }
    }
    if (RC) { GRcAcc acc = {s, L}; feat_cell(acc, 16 * (i64)c, lo, hi); }
    else { GAcc acc = {s, L}; feat_cell(acc, 16 * (i64)c, lo, hi); }
}

// genome: aligned 16-byte loads for interior cells
struct GAcc16
{
    const u8 * s; i64 len;
    __device__ __forceinline__ int operator()(i64 p) const { return (p >= 0 && p < len) ? (int)__ldg(s + p) : 0; }
};

__global__ void __launch_bounds__(FT) k_feat_genome(const u8 * __restrict__ g, i64 len, u32 n_entries, F96 * __restrict__ out)
{
    __shared__ u64 s_lo[FT];
    __shared__ u32 s_hi[FT];
    u32 e0 = blockIdx.x * FE;
    u32 c = e0 + threadIdx.x;
    u64 lo = 0; u32 hi = 0;
    i64 p0 = 16 * (i64)c;
    if (c < n_entries + 2)
    {
        if (p0 + 17 <= len)
        {
            // one aligned 16-byte load + the first byte of the next cell
            uint4 v = __ldg((const uint4 *)(g + p0));
            u32 wds[4] = {v.x, v.y, v.z, v.w};
            int a = (int)(wds[0] & 0xff);
#pragma unroll
            for (int i = 1; i <= 16; i++)
            {
                int b = i < 16 ? (int)((wds[i >> 2] >> (8 * (i & 3))) & 0xff) : (int)__ldg(g + p0 + 16);
                if (a < 4 && b < 4)
                {
                    int id = 4 * a + b;
                    if (id < 10) lo += 1ULL << (6 * id);
                    else if (id < 15) hi += 1u << (6 * (id - 10));
                }
                a = b;
            }
        }
        else
        {
            GAcc acc = {g, len};
            feat_cell(acc, p0, lo, hi);
        }
    }
    s_lo[threadIdx.x] = lo;
    s_hi[threadIdx.x] = hi;
    __syncthreads();
    if (threadIdx.x < FE && c < n_entries)
    {
        F96 f = feat_entry(lo + s_lo[threadIdx.x + 1] + s_lo[threadIdx.x + 2], hi + s_hi[threadIdx.x + 1] + s_hi[threadIdx.x + 2]);
        i32 * o = (i32 *)(out + c);
        o[0] = f.v[0]; o[1] = f.v[1]; o[2] = f.v[2];
    }
}

// reads: ftile[i] = first tile of read i (2 strands per read); tile_read[t] = the read tile t belongs to
__global__ void k_feat_tile_reads(const u32 * __restrict__ ftile, u32 n_reads, u32 * __restrict__ tile_read)
{
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    for (u32 t = ftile[r]; t < ftile[r + 1]; t++) tile_read[t] = r;
}
// one descriptor per pair of tiles (forward + reverse tile of the same index), laid out by a thread per read: the feature
// kernel then starts an iteration with ONE load instead of a chain tile -> read -> offsets
struct FeatPair { u64 base; u64 out; u32 L; u32 nf; u32 e0; u32 pad; };
__global__ void k_feat_pairs(const u32 * __restrict__ ftile, const u64 * __restrict__ read_off, const u64 * __restrict__ foff, u32 n_reads,
                             FeatPair * __restrict__ pairs)
{
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    FeatPair d;
    d.base = read_off[r]; d.out = foff[r];
    const u64 L = read_off[r + 1] - read_off[r];
    d.L = (u32)L; d.nf = feat_count_read(L); d.pad = 0;
    const u32 p0 = ftile[r] >> 1, p1 = ftile[r + 1] >> 1;
    for (u32 p = p0; p < p1; p++) { d.e0 = (p - p0) * FE; pairs[p] = d; }
}
// persistent CTAs over the tiles: the 2-mer table is built once per CTA, not once per 256 cells. A read's tiles come as
// tps forward tiles followed by tps reverse tiles (n_tiles is even, every read's range starts at an even tile): one loop
// iteration takes the forward and the reverse tile of the same index together -- both cells' loads are issued before either
// is used. LNR_FEAT_BUFS = 2 exchanges the cell sums through two alternating shared buffers, so that an iteration has ONE
// barrier (a thread can be at most one iteration ahead of the slowest, and then writes the other buffer).
__global__ void __launch_bounds__(FT, 8) k_feat_reads(const u8 * __restrict__ bases, const FeatPair * __restrict__ pairs, u32 n_tiles,
                                                   F96 * __restrict__ out)
{
#ifndef LNR_FEAT_BUFS
#define LNR_FEAT_BUFS 1     // 2: alternating buffers, one barrier per iteration -- measured equal (1.57 vs 1.60 ms, call 29), 6 KB more
#endif
    __shared__ u64 s_lo[LNR_FEAT_BUFS][2][FT];
    __shared__ u32 s_hi[LNR_FEAT_BUFS][2][FT];
    __shared__ FeatTab T;
    feat_tab_init(T);
    __syncthreads();
    const u32 n_pairs = n_tiles >> 1;
    u32 buf = 0;
    for (u32 pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, buf ^= (u32)(LNR_FEAT_BUFS - 1))
    {
        const FeatPair d = pairs[pair];
        const u64 L = d.L;
        const u32 nf = d.nf;
        const u32 c = d.e0 + threadIdx.x;
        F96 * o = out + d.out;
        const u8 * s = bases + d.base;
        CellLoad cf, cr;
        feat_cell_issue<false>(bases, s, (i64)L, c, nf, cf);
        feat_cell_issue<true>(bases, s, (i64)L, c, nf, cr);
        u64 lo0, lo1; u32 hi0, hi1;
        feat_cell_finish<false>(cf, s, (i64)L, c, T, lo0, hi0);
        feat_cell_finish<true>(cr, s, (i64)L, c, T, lo1, hi1);
        u64 (* bl)[FT] = s_lo[buf]; u32 (* bh)[FT] = s_hi[buf];
        bl[0][threadIdx.x] = lo0; bh[0][threadIdx.x] = hi0;
        bl[1][threadIdx.x] = lo1; bh[1][threadIdx.x] = hi1;
        __syncthreads();
        if (threadIdx.x < FE && c < nf)
        {
            F96 f = feat_entry(lo0 + bl[0][threadIdx.x + 1] + bl[0][threadIdx.x + 2], hi0 + bh[0][threadIdx.x + 1] + bh[0][threadIdx.x + 2]);
            i32 * w = (i32 *)(o + c);
            w[0] = f.v[0]; w[1] = f.v[1]; w[2] = f.v[2];
            f = feat_entry(lo1 + bl[1][threadIdx.x + 1] + bl[1][threadIdx.x + 2], hi1 + bh[1][threadIdx.x + 1] + bh[1][threadIdx.x + 2]);
            w = (i32 *)(o + nf + c);
            w[0] = f.v[0]; w[1] = f.v[1]; w[2] = f.v[2];
        }
#if LNR_FEAT_BUFS == 1
        __syncthreads();
#endif
    }
}

// ---- -f 1: 1-mer / 32-base features (createFeatures1_32 pmpfinder.cpp:354 serial / :393 parallel) --------------------
// entry i = A + 32 C + 1024 G counts over the 32 bases from 16 i on, as a short (a count of 32 carries, the sum wraps:
// part of the spec). One thread per 16-base cell, an entry is the sum of two neighbouring cells (through shared memory).
// `written` = entries the reference's builder writes; the rest of the string is 0 (canonical rule).
template <bool RC>
__device__ __forceinline__ int feat32_cell(const u8 * __restrict__ s, i64 L, i64 c)   // cell c of the (reverse-complemented) sequence
{
    int v = 0;
#pragma unroll
    for (int k = 0; k < 16; k++)
    {
        const i64 p = 16 * c + k;
        int b = 4;
        if (p < L) { b = (int)__ldg(s + (RC ? L - 1 - p : p)); if (RC && b < 4) b = 3 - b; }
        v += b == 0 ? 1 : (b == 1 ? 32 : (b == 2 ? 1024 : 0));
    }
    return v;
}
__global__ void __launch_bounds__(FT) k_feat32_genome(const u8 * __restrict__ g, i64 len, u32 n_written, i16 * __restrict__ out)
{
    __shared__ int s_c[FT];
    const u32 e0 = blockIdx.x * (FT - 1);
    const u32 c = e0 + threadIdx.x;
    s_c[threadIdx.x] = c < n_written + 1 ? feat32_cell<false>(g, len, (i64)c) : 0;
    __syncthreads();
    if (threadIdx.x < FT - 1 && c < n_written) out[c] = (i16)(s_c[threadIdx.x] + s_c[threadIdx.x + 1]);
}
// reads: grid.x = tiles of (FT - 1) entries, grid.y = (read, strand); both strands' strings back to back at foff[r]
__global__ void __launch_bounds__(FT) k_feat32_reads(const u8 * __restrict__ bases, const u64 * __restrict__ read_off, const u64 * __restrict__ foff,
                                                    const u32 * __restrict__ ftile, const u32 * __restrict__ tile_read, u32 n_tiles, i16 * __restrict__ out)
{
    __shared__ int s_c[FT];
    for (u32 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
    {
        const u32 r = tile_read[tile];
        const u64 L = read_off[r + 1] - read_off[r];
        const u32 nf = feat32_count(L), nw = feat32_written_serial(L);
        const u32 tps = (nf + (FT - 1) - 1) / (FT - 1);
        const u32 t = tile - ftile[r];
        const u32 strand = t >= tps ? 1u : 0u;
        const u32 c = (t - strand * tps) * (FT - 1) + threadIdx.x;
        const u8 * s = bases + read_off[r];
        int v = 0;
        if (c < nw + 1) v = strand ? feat32_cell<true>(s, (i64)L, (i64)c) : feat32_cell<false>(s, (i64)L, (i64)c);
        s_c[threadIdx.x] = v;
        __syncthreads();
        i16 * o = out + foff[r] + (u64)strand * nf;
        if (threadIdx.x < FT - 1 && c < nf) o[c] = c < nw ? (i16)(s_c[threadIdx.x] + s_c[threadIdx.x + 1]) : (i16)0;
        __syncthreads();
    }
}

// =====================================================================================================
// DIndex build
// =====================================================================================================
static const int IT = 256;           // threads per CTA
static const int IS = 4;             // samples per thread
static const int ITILE = IT * IS;    // samples per CTA
static const int ISM = ITILE * 9 + 64 + 32;

__global__ void k_idx_prep(const u8 * __restrict__ g, IdxChunk * chunks, u32 n_chunks)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chunks) return;
    IdxChunk ch = chunks[i];
    GAcc acc = {g + ch.base_off, ch.len};
    ch.kskip = hash_init_skip<kSpanD>(acc, ch.t_str, ch.len);
    ch.bias = ch.kskip ? selector_bias<kSpanD>(acc, ch.t_str + ch.kskip, ch.t_str) : 0;
    chunks[i] = ch;
}

struct TileAcc   // smem-staged tile with global fallback
{
    const u8 * sm; i64 p0, p1; const u8 * g; i64 len;
    __device__ __forceinline__ int operator()(i64 p) const
    {
        if (p >= p0 && p < p1) return (int)sm[p - p0];
        return (p >= 0 && p < len) ? (int)__ldg(g + p) : 0;
    }
};

// ---- device-wide exclusive scan of u32 (reduce / spine / apply), 4096 items per CTA --------------------
static const int ST = 256, SI = 16, STILE = ST * SI;
// transform applied to the input: cap > 0 -> values > cap become 0 and are written back (bucket omission)
__global__ void __launch_bounds__(ST) k_scan_reduce(u32 * __restrict__ in, u64 n, u32 cap, u64 * __restrict__ partial)
{
    __shared__ u64 s[ST / 32];
    u64 base = (u64)blockIdx.x * STILE;
    u64 sum = 0;
    for (int i = 0; i < SI; i++)
    {
        u64 idx = base + (u64)i * ST + threadIdx.x;
        if (idx < n)
        {
            u32 v = in[idx];
            if (cap && v > cap) { v = 0; in[idx] = 0; }
            sum += v;
        }
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) { u64 t = 0; for (int i = 0; i < ST / 32; i++) t += s[i]; partial[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024) k_scan_spine(u64 * __restrict__ partial, u32 n_blocks, u64 * __restrict__ total)
{
    __shared__ u64 s[32];
    __shared__ u64 s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (u32 base = 0; base < n_blocks; base += 1024)
    {
        u32 i = base + threadIdx.x;
        u64 v = i < n_blocks ? partial[i] : 0;
        u64 x = v;
        for (int o = 1; o < 32; o <<= 1) { u64 t = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += t; }
        if ((threadIdx.x & 31) == 31) s[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32)
        {
            u64 y = s[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) { u64 t = __shfl_up_sync(0xffffffffu, y, o); if (threadIdx.x >= o) y += t; }
            s[threadIdx.x] = y;
        }
        __syncthreads();
        u64 wbase = (threadIdx.x >> 5) ? s[(threadIdx.x >> 5) - 1] : 0;
        u64 incl = x + wbase + s_carry;
        if (i < n_blocks) partial[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}
// OutT = i32 (dir) or u32 / u64 offsets
// A warp owns 512 consecutive items as 4 chunks of 128 and a lane 4 consecutive items of every chunk, so that a warp's loads
// are contiguous 512-byte runs and its stores contiguous 0.5 / 1-KB runs (16 consecutive items per thread -- the first
// version -- made every request touch 32 sectors: the pass ran at 1.9 TB/s).
template <class OutT>
__global__ void __launch_bounds__(ST) k_scan_apply(const u32 * __restrict__ in, u64 n, const u64 * __restrict__ partial, OutT * __restrict__ out)
{
    static_assert(STILE == (ST / 32) * 512, "a warp scans 512 items");
    __shared__ u64 s[ST / 32];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const u64 wb = (u64)blockIdx.x * STILE + (u64)wid * 512;
    const bool vin = ((uintptr_t)in & 15) == 0, vout = ((uintptr_t)out & 15) == 0;
    u32 v[4][4];
    u64 lsum[4], incl[4], ctot[4];
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
        const u64 idx = wb + 128u * j + 4u * lane;
        if (vin && idx + 3 < n)
        {
            const uint4 q = *(const uint4 *)(in + idx);
            v[j][0] = q.x; v[j][1] = q.y; v[j][2] = q.z; v[j][3] = q.w;
        }
        else
        {
#pragma unroll
            for (int k = 0; k < 4; k++) v[j][k] = idx + k < n ? in[idx + k] : 0u;
        }
        lsum[j] = (u64)v[j][0] + v[j][1] + v[j][2] + v[j][3];
    }
    u64 wtot = 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
        u64 x = lsum[j];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u64 t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (unsigned)o) x += t; }
        incl[j] = x;
        ctot[j] = __shfl_sync(0xffffffffu, x, 31);
        wtot += ctot[j];
    }
    if (lane == 0) s[wid] = wtot;
    __syncthreads();
    u64 carry = partial[blockIdx.x];
    for (unsigned i = 0; i < wid; i++) carry += s[i];
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
        const u64 idx = wb + 128u * j + 4u * lane;
        const u64 r0 = carry + incl[j] - lsum[j], r1 = r0 + v[j][0], r2 = r1 + v[j][1], r3 = r2 + v[j][2];
        if (vout && idx + 3 < n)
        {
            if (sizeof(OutT) == 8)
            {
                ulonglong2 a, b2;
                a.x = r0; a.y = r1; b2.x = r2; b2.y = r3;
                ((ulonglong2 *)(out + idx))[0] = a; ((ulonglong2 *)(out + idx))[1] = b2;
            }
            else
            {
                uint4 q;
                q.x = (u32)(OutT)r0; q.y = (u32)(OutT)r1; q.z = (u32)(OutT)r2; q.w = (u32)(OutT)r3;
                *(uint4 *)(out + idx) = q;
            }
        }
        else
        {
            if (idx < n) out[idx] = (OutT)r0;
            if (idx + 1 < n) out[idx + 1] = (OutT)r1;
            if (idx + 2 < n) out[idx + 2] = (OutT)r2;
            if (idx + 3 < n) out[idx + 3] = (OutT)r3;
        }
        carry += ctot[j];
    }
}

// SoA split for the seeding count pass: the Y-key rule only looks at the low byte of a record (Y is 8 bits,
// shape_extend.cpp:282), so buckets are scanned as bytes (1/8 of the traffic) and hs itself is read only for matches
__global__ void k_idx_split_y(const u64 * __restrict__ hs, u64 n, u8 * __restrict__ hsy)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) hsy[i] = (u8)(hs[i] & 0xff);
}

// Lookup entry of the seeding count pass: dir[X], the bucket size and the bucket's first 56 Y keys in 64 bytes = exactly what
// one L2 miss brings in when the load carries the 64-byte fetch hint (ld_dirx below; a plain load fetches 128 bytes). A seed
// whose bucket has <= 56 records (95 % of them at 3.1 Gbase) costs one random DRAM access; with 32-byte entries (24 keys)
// every second seed paid a second, dependent one into hsy.
static const int kDirxKeys = 56, kDirxQuads = 4;      // keys per entry, uint4 per entry
__global__ void k_idx_dirx(const i32 * __restrict__ dir, const u8 * __restrict__ hsy, u32 n_buckets, uint4 * __restrict__ out)
{
    u32 X = blockIdx.x * blockDim.x + threadIdx.x;
    if (X >= n_buckets) return;
    const i32 b = dir[X], e = dir[X + 1];
    const u32 n = (u32)(e - b);
    u32 wds[14];
#pragma unroll
    for (int q = 0; q < 14; q++)
    {
        u32 v = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { u32 idx = 4 * q + k; if (idx < n) v |= (u32)hsy[(size_t)b + idx] << (8 * k); }
        wds[q] = v;
    }
    uint4 * o = out + (size_t)kDirxQuads * X;
    o[0] = make_uint4((u32)b, n, wds[0], wds[1]);
    o[1] = make_uint4(wds[2], wds[3], wds[4], wds[5]);
    o[2] = make_uint4(wds[6], wds[7], wds[8], wds[9]);
    o[3] = make_uint4(wds[10], wds[11], wds[12], wds[13]);
}

// =====================================================================================================
// seeding (getDIndexMatchAll, pmpfinder.cpp:1856)
// =====================================================================================================
// Besides the hash constants of a task this pass lays out what the two seeding kernels would otherwise find through
// chains of dependent loads at the head of every warp: the read's offset and length next to the task (tinfo), and for
// every warp of 32 samples the task of its first sample (wrec[].task; the count pass rewrites the record with the same
// value) -- a binary search over 65 536 tasks was 16 round trips before a warp touched its first base.
struct SeedTaskInfo { u64 base; u64 len; };
struct SeedWarpRec { u32 list_off; u32 scanned; u32 task; u32 pad; };   // per warp: start of its entries (0xffffffff: none), records scanned (H), task of its first sample
// Random index reads of the seeding kernels. A plain load that misses in L2 brings in 128 bytes (measured: it times exactly
// like ld.global.nc.L2::128B, and ncu shows 2.8 DRAM sectors per requested sector); the smallest size PTX can ask for is 64
// bytes -- one lookup entry of dirx, or the sector pair around an 8-byte hs record: k_seed_fill 2.49 -> 2.17 ms per 65 536
// reads (L2::256B: 2.67 ms; L2-only ld.global.cg loads and a 32-byte cudaLimitMaxL2FetchGranularity change nothing).
__device__ __forceinline__ u64 ld_hs(const u64 * p)
{
    u64 v; asm volatile("ld.global.nc.L2::64B.b64 %0, [%1];" : "=l"(v) : "l"(p)); return v;
}
__device__ __forceinline__ uint4 ld_dirx(const uint4 * p)
{
#ifdef LNR_DIRX_PLAIN
    return __ldg(p);
#else
    uint4 v; asm volatile("ld.global.nc.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v;
#endif
}
__global__ void k_seed_prep(const u8 * __restrict__ bases, const u64 * __restrict__ read_off, SeedTask * tasks, u32 n_tasks,
                            SeedTaskInfo * __restrict__ tinfo, SeedWarpRec * __restrict__ wrec)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tasks) return;
    SeedTask t = tasks[i];
    u64 L = read_off[t.read + 1] - read_off[t.read];
    GAcc acc = {bases + read_off[t.read], (i64)L};
    t.kskip = (u32)hash_init_skip<kSpanD>(acc, 0, (i64)L);
    t.bias = selector_bias<kSpanD>(acc, (i64)t.kskip, (i64)t.str + kSpanD);
    tasks[i] = t;
    SeedTaskInfo inf; inf.base = read_off[t.read]; inf.len = L;
    tinfo[i] = inf;
    // warps w whose first sample 32 w lies in [sample0, sample0 + n_samples)
    const u64 w1 = (t.sample0 + t.n_samples + 31) >> 5;
    for (u64 w = (t.sample0 + 31) >> 5; w < w1; w++) wrec[w].task = i;
}

// per sample: packed info = bucket start (32) | bucket size (16) | Y (8) | strand (1) | k (is recomputed) ;
// count = records passing the Y-key rule. A sample whose X equals the previous sample's X is skipped
// (xpre rule, :1882; xpre starts at 0).
__device__ __forceinline__ u32 find_task(const SeedTask * tasks, u32 n_tasks, u64 s)
{
    u32 lo = 0, hi = n_tasks;
    while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (tasks[mid].sample0 <= s) lo = mid; else hi = mid; }
    return lo;
}

// Fast evaluation of one pure 21-base window (>= span steps after hashInit) whose 4 bases of context on either side are
// readable and free of N: the 32 bytes around it are fetched as 8 aligned words (global memory or a staged shared-memory
// tile), packed to 2 bits each (one multiply per 4 bases), and the forward hash, the reverse-complement hash (bit
// reversal), the selector sum (popcounts), the leftmost minimal 13-mer (min over offset-tagged keys) and both flank
// keys are derived from that 58-bit value -- ~120 instead of ~500 instructions. q points at window start - 4; the words
// [q & ~3, (q & ~3) + 32) must be readable. Returns false when one of the 29 bases is an N.
template <bool SMEM>
__device__ __forceinline__ bool window_fast(const u8 * q, int bias, SeedVal & sv)
{
    const unsigned sh = ((unsigned)(uintptr_t)q & 3u) * 8u;
    const u32 * pw = (const u32 *)((uintptr_t)q & ~(uintptr_t)3);
    u32 w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = SMEM ? pw[i] : __ldg(pw + i);
    u32 A[8], any = 0;
#pragma unroll
    for (int i = 0; i < 7; i++) A[i] = __funnelshift_r(w[i], w[i + 1], sh);
    A[7] = (w[7] >> sh) & 0xffu;
#pragma unroll
    for (int i = 0; i < 8; i++) any |= A[i];
    if (any & 0xfcfcfcfcu) return false;
    u32 qb[8];
#pragma unroll
    for (int i = 0; i < 8; i++) qb[i] = (A[i] * 0x40100401u) >> 24;   // bases 4i..4i+3 -> 8 bits, first base on top
    const u32 Phi = (qb[0] << 24) | (qb[1] << 16) | (qb[2] << 8) | qb[3];
    const u32 Plo = (qb[4] << 24) | (qb[5] << 16) | (qb[6] << 8) | qb[7];
    const u64 P = ((u64)Phi << 32) | Plo;              // base j (position p-4+j) at bits 2(31-j)
    const u64 M42 = (1ULL << 42) - 1;
    const u64 h = (P >> 14) & M42;                      // window = j 4..24
    // complement + reverse the pairs: base j at bits 2j
    u64 y = __brevll(~P);
    const u64 Pc = ((y >> 1) & 0x5555555555555555ULL) | ((y & 0x5555555555555555ULL) << 1);
    const u64 cr = (Pc >> 8) & M42;
    const int sum = __popcll(h & 0x5555555555555555ULL) + 2 * __popcll(h & 0xAAAAAAAAAAAAAAAAULL);
    const int x = 2 * sum - 3 * kSpanD + bias;
    const u32 strand = x > 0 ? 0u : 1u;
    const u64 v2 = strand ? cr : h;
    u32 best = 0xffffffffu;
#pragma unroll
    for (int o = 0; o <= 8; o++)
    {
        const int s2 = 2 * (8 - o) - 4;                 // minimizer candidate into bits 4..29, offset tag below it
        u32 key = s2 >= 0 ? (u32)(v2 >> s2) : (u32)(v2 << (-s2));
        key = (key & 0x3ffffff0u) | (u32)o;
        best = min(best, key);
    }
    const u32 off = best & 15u;
    sv.X = best >> 4;
    sv.strand = strand;
    sv.Y = strand ? (u32)(Pc >> (2 * (8 - off))) & 0xffu : (u32)(P >> (2 * (11 - off))) & 0xffu;
    return true;
}
// read seeding: sample m of a task (seed_sample, lnr_core.h); false = the sample needs the general evaluation
__device__ __forceinline__ bool seed_sample_fast(const u8 * __restrict__ rd, i64 L, const SeedTask & t, u32 m, SeedVal & sv)
{
    const i64 k0 = (i64)t.str + kSpanD;
    const i64 p = k0 + (i64)t.alpha * m - 1;
    if (p - k0 + 1 < kSpanD || p < 8 || p + 28 > L) return false;
    return window_fast<false>(rd + p - 4, t.bias, sv);
}
// =====================================================================================================
// DIndex build, second design (round 2): hash ONCE, partition, then count / place / sort inside L2.
//   k_idx_emit       CTA = 1024 consecutive samples of a chunk; the 9.3 KB genome tile arrives in shared memory as one
//                    TMA bulk copy (cp.async.bulk + mbarrier); samples take the packed fast path from the tile; emit rule
//                    by block max-scan; every sample slot gets (X | invalid, record) -- dense, 16/32-byte stores. Every
//                    32nd tile also feeds a 4096-bin histogram of X >> 14, from which the host picks 64 splitters of
//                    ~equal record count (the minimizer distribution is skewed towards small X and genome dependent)
//   k_idx_partcount  exact number of pairs per partition
//   k_idx_part       pairs -> the 64 partitions, staged per CTA in shared memory so that every partition receives
//                    contiguous runs
//   k_idx_count      cnt[X]++ over the partitioned pairs: one launch, CTAs walk the partitions in order, so the atomics
//                    of the CTAs in flight fall into a window of cnt that L2 holds (the first design's histogram pass
//                    dirtied one 32-byte sector per sample: 11 GB of write-backs)
//   k_scan_*         bucket omission + exclusive scan -> dir (unchanged)
//   k_idx_place      per partition: record -> staging[dir[X] + slot]; the partition's range (~43 MB) stays in L2, so the
//                    random 8-byte stores combine there instead of each costing a DRAM read-modify-write
//   k_idx_rank       per partition, right behind its place kernel (L2-hot): every record counts the smaller records of
//                    its bucket and moves to that rank of hs (ascending order inside a bucket), its Y byte into hsy
//   k_idx_dirx       the lookup sectors of the seeding pass, one streaming pass at the end
// The first design (two hashing passes + atomic scatter + one thread-per-bucket sort from DRAM) moved 70 GB of DRAM traffic
// in its scatter alone (profiles/r2_ncu_idx_full_summary.txt).
// =====================================================================================================
static const int IPARTS = 64;
static const int ICOARSE_SHIFT = 14, ICOARSE = 1 << (kDirBits - ICOARSE_SHIFT);   // 4096 coarse bins for the splitters
struct IdxParts { u32 split[IPARTS + 1]; };   // partition p holds X in [split[p], split[p+1])

__device__ __forceinline__ u32 smem_u32(const void * p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 * bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d(void * dst_smem, const void * src_gmem, u32 bytes, u64 * bar)
{
    // one thread: arm the barrier with the byte count, then one bulk async copy global -> shared (TMA, UBLKCP)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 * bar, u32 parity)
{
    asm volatile("{\n .reg .pred p;\n LNR_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra LNR_DONE;\n bra LNR_WAIT;\n LNR_DONE:\n}" ::"r"(
                     smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ u32 idx_part_of(const u32 * split, u32 X)
{
    u32 lo = 0;
#pragma unroll
    for (int step = IPARTS / 2; step; step >>= 1)
        if (split[lo + step] <= X) lo += step;
    return lo;
}

// the general evaluation (chunk starts, N, contig ends) is rare: one out-of-line copy keeps the kernel's hot code small
__device__ __noinline__ void idx_sample_slow(const u8 * gs, const IdxChunk & ch, i64 m, u32 & X, u64 & rec)
{
    GAcc ga = {gs, ch.len};
    idx_sample(ga, ch, m, X, rec);
}
// sample m of a chunk from the staged tile (fast path) or through the bounds-checked accessor
__device__ __forceinline__ void idx_sample_tile(const u8 * s_b, i64 p0, i64 p1, const TileAcc & acc, const IdxChunk & ch, i64 m, u32 & X, u64 & rec)
{
    const i64 j = ch.t_str + kIdxMinStep + 9 * m;
    SeedVal sv;
    if (j - ch.t_str + 1 >= kSpanD && j - 4 >= p0 && j + 28 <= p1 && window_fast<true>(s_b + (j - 4 - p0), ch.bias, sv))
    {
        X = sv.X;
        rec = create_cord(ch.contig, (u64)j + kAnchorZero, sv.Y, sv.strand);
        return;
    }
    idx_sample_slow(acc.g, ch, m, X, rec);
}
__device__ __forceinline__ u32 idx_sample_x_global(const u8 * gs, const IdxChunk & ch, i64 m)
{
    const i64 j = ch.t_str + kIdxMinStep + 9 * m;
    SeedVal sv;
    if (j - ch.t_str + 1 >= kSpanD && j >= 8 && j + 28 <= ch.len && window_fast<false>(gs + j - 4, ch.bias, sv)) return sv.X;
    u32 X; u64 r;
    idx_sample_slow(gs, ch, m, X, r);
    return X;
}

__global__ void __launch_bounds__(IT) k_idx_emit(const u8 * __restrict__ g, const IdxChunk * __restrict__ chunks,
                                                 const u32 * __restrict__ tile0, u32 n_chunks, u32 * __restrict__ pairX,
                                                 u64 * __restrict__ pairRec, u32 x_lo, u32 x_hi, u32 * __restrict__ coarse)
{
    // [x_lo, x_hi): minimizer range owned by this shard (multi-GPU build partitions the 2^26 buckets)
    __shared__ __align__(128) u8 s_b[(ISM + 15 + 16) & ~15];
    __shared__ __align__(8) u64 s_bar;
    __shared__ i32 s_warp[IT / 32];
    __shared__ u32 s_last[IT / 32];
    __shared__ i32 s_back;
    const u32 tile = blockIdx.x;
    u32 lo = 0, hi = n_chunks;
    while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (tile0[mid] <= tile) lo = mid; else hi = mid; }
    const IdxChunk ch = chunks[lo];
    const u8 * gs = g + ch.base_off;
    const i64 m0 = (i64)(tile - tile0[lo]) * ITILE;
    const i64 nm = ch.n_samples - m0 < ITILE ? ch.n_samples - m0 : ITILE;
    // stage bases [p0, p1): p0 16-byte aligned (contigs start at 256-byte aligned offsets), the copy length rounded up to
    // 16 -- at most 15 bytes into the >= 256 zero bytes behind the contig
    const i64 j0 = ch.t_str + kIdxMinStep + 9 * m0;
    i64 p0 = (j0 - 8) & ~15LL;
    if (p0 < 0) p0 = 0;
    i64 p1 = j0 + 9 * nm + 32;
    if (p1 > ch.len) p1 = ch.len;
    const u32 nbytes = (u32)((p1 - p0 + 15) & ~15LL);
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) tma_load_1d(s_b, gs + p0, nbytes, &s_bar);
    // while the tile is in flight: look-back of the tile's first sample (consecutive earlier samples with the same X);
    // it needs X of sample m0, which lane 0 takes from global memory like the earlier ones
    if (threadIdx.x < 32)
    {
        i32 back = 0;
        if (m0 > 0)
        {
            u32 X0 = 0;
            if (threadIdx.x == 0) X0 = idx_sample_x_global(gs, ch, m0);
            X0 = __shfl_sync(0xffffffffu, X0, 0);
            i64 m = m0 - 1;
            while (true)
            {
                i64 mm = m - threadIdx.x;
                bool same = false;
                if (mm >= 0) same = idx_sample_x_global(gs, ch, mm) == X0;
                u32 neq = __ballot_sync(0xffffffffu, !same);
                if (neq) { back += __ffs((int)neq) - 1; break; }
                back += 32; m -= 32;
            }
        }
        if (threadIdx.x == 0) s_back = back;
    }
    mbar_wait(&s_bar, 0);
    TileAcc acc = {s_b, p0, p1, gs, ch.len};
    // each thread: IS consecutive samples
    u32 X[IS]; u64 rec[IS];
    const i64 mt = m0 + (i64)threadIdx.x * IS;
#pragma unroll
    for (int q = 0; q < IS; q++)
    {
        X[q] = 0xffffffffu; rec[q] = 0;
        if ((i64)threadIdx.x * IS + q < nm) idx_sample_tile(s_b, p0, p1, acc, ch, mt + q, X[q], rec[q]);
    }
    // X of the sample just before this thread's first one
    u32 xprev = __shfl_up_sync(0xffffffffu, X[IS - 1], 1);
    if ((threadIdx.x & 31) == 31) s_last[threadIdx.x >> 5] = X[IS - 1];
    __syncthreads();
    if ((threadIdx.x & 31) == 0) xprev = threadIdx.x == 0 ? 0xfffffffeu : s_last[(threadIdx.x >> 5) - 1];
    // run start (tile-relative sample index; negative = before the tile) by max-scan
    const i32 local_i = (i32)threadIdx.x * IS;
    i32 rs[IS];
    i32 run = -0x40000000;   // "unknown, inherited"
    {
        u32 xp = xprev;
#pragma unroll
        for (int q = 0; q < IS; q++)
        {
            bool brk = X[q] != xp;
            if (threadIdx.x == 0 && q == 0) brk = false;   // resolved through s_back
            if (brk) run = local_i + q;
            rs[q] = run;
            xp = X[q];
        }
    }
    i32 v = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { i32 t = __shfl_up_sync(0xffffffffu, v, o); if ((threadIdx.x & 31) >= o) v = max(v, t); }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    i32 carry = -0x40000000;
    for (int wq = 0; wq < (int)(threadIdx.x >> 5); wq++) carry = max(carry, s_warp[wq]);
    i32 excl = __shfl_up_sync(0xffffffffu, v, 1);
    if ((threadIdx.x & 31) == 0) excl = -0x40000000;
    excl = max(excl, carry);
    const i32 tile_start_run = -s_back;   // run start of sample 0 (<= 0)
    u32 ox[IS];
#pragma unroll
    for (int q = 0; q < IS; q++)
    {
        const i32 r = rs[q] > -0x40000000 ? rs[q] : (excl > -0x40000000 ? excl : tile_start_run);
        const i32 idx = local_i + q;
        const bool emit = idx < nm && (((idx - r) & 1) == 0) && X[q] >= x_lo && X[q] < x_hi;
        ox[q] = emit ? X[q] : 0xffffffffu;
        if (emit && (tile & 31u) == 0) atomicAdd(&coarse[X[q] >> ICOARSE_SHIFT], 1u);   // a 1/32 sample is plenty for the splitters
    }
    // the tile's slots: 16 bytes of X and 32 bytes of records per thread
    const u64 slot = (u64)tile * ITILE + (u64)local_i;
    *(uint4 *)(pairX + slot) = make_uint4(ox[0], ox[1], ox[2], ox[3]);
    *(ulonglong2 *)(pairRec + slot) = make_ulonglong2(rec[0], rec[1]);
    *(ulonglong2 *)(pairRec + slot + 2) = make_ulonglong2(rec[2], rec[3]);
}

// exact number of pairs per partition (one pass over the X column)
__global__ void __launch_bounds__(256) k_idx_partcount(const u32 * __restrict__ pairX, u64 n_slots, IdxParts parts, unsigned long long * __restrict__ part_cnt,
                                                       u32 x_lo, u32 x_hi)
{
    __shared__ u32 s_split[IPARTS + 1];
    __shared__ u32 s_pc[IPARTS];
    if (threadIdx.x < IPARTS + 1) s_split[threadIdx.x] = parts.split[threadIdx.x];
    if (threadIdx.x < IPARTS) s_pc[threadIdx.x] = 0;
    __syncthreads();
    const u64 n4 = n_slots / 4;     // n_slots is a multiple of the tile size
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (u64)gridDim.x * blockDim.x)
    {
        const uint4 v = __ldg((const uint4 *)pairX + i);
        const u32 xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (xs[q] >= x_lo && xs[q] < x_hi) atomicAdd(&s_pc[idx_part_of(s_split, xs[q])], 1u);   // 0xffffffff (no record) is never < x_hi
    }
    __syncthreads();
    if (threadIdx.x < IPARTS && s_pc[threadIdx.x]) atomicAdd(&part_cnt[threadIdx.x], (unsigned long long)s_pc[threadIdx.x]);
}
// histogram of the partitioned pairs (see the header comment); 4 independent atomics per thread: the kernel is latency bound
__global__ void __launch_bounds__(256) k_idx_count(const u32 * __restrict__ X, u64 n, u32 * __restrict__ cnt)
{
    const u64 i0 = (u64)blockIdx.x * (blockDim.x * 4) + threadIdx.x;
    u32 x[4];
#pragma unroll
    for (int q = 0; q < 4; q++) { const u64 i = i0 + (u64)q * blockDim.x; x[q] = i < n ? __ldg(X + i) : 0xffffffffu; }
#pragma unroll
    for (int q = 0; q < 4; q++) if (x[q] != 0xffffffffu) atomicAdd(&cnt[x[q]], 1u);
}
static const int PT = 256, PI = 8, PTILE = PT * PI;
__global__ void __launch_bounds__(PT) k_idx_part(const u32 * __restrict__ pairX, const u64 * __restrict__ pairRec, u64 n_slots, IdxParts parts,
                                                 const u64 * __restrict__ part_off, unsigned long long * __restrict__ part_fill,
                                                 u32 * __restrict__ outX, u64 * __restrict__ outRec, u32 x_lo, u32 x_hi)
{
    __shared__ u32 s_split[IPARTS + 1];
    __shared__ u32 s_cnt[IPARTS], s_start[IPARTS + 1];
    __shared__ u64 s_gbase[IPARTS];
    __shared__ u32 sX[PTILE];
    __shared__ u64 sRec[PTILE];
    __shared__ u8 sP[PTILE];
    if (threadIdx.x < IPARTS + 1) s_split[threadIdx.x] = parts.split[threadIdx.x];
    if (threadIdx.x < IPARTS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const u64 base = (u64)blockIdx.x * PTILE;
    u32 x[PI], rk[PI], pp[PI]; u64 rc[PI];
#pragma unroll
    for (int i = 0; i < PI; i++)
    {
        const u64 idx = base + (u64)i * PT + threadIdx.x;
        x[i] = idx < n_slots ? __ldg(pairX + idx) : 0xffffffffu;
        pp[i] = 0xffu; rk[i] = 0; rc[i] = 0;
        if (x[i] >= x_lo && x[i] < x_hi)
        {
            rc[i] = __ldg(pairRec + idx);
            pp[i] = idx_part_of(s_split, x[i]);
            rk[i] = atomicAdd(&s_cnt[pp[i]], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < 32)
    {
        // exclusive scan of the 64 counts (2 per lane), then one global claim per non-empty partition
        const u32 c0 = s_cnt[2 * threadIdx.x], c1 = s_cnt[2 * threadIdx.x + 1];
        u32 incl = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (threadIdx.x >= (unsigned)o) incl += t; }
        const u32 ex = incl - c0 - c1;
        s_start[2 * threadIdx.x] = ex; s_start[2 * threadIdx.x + 1] = ex + c0;
        if (threadIdx.x == 31) s_start[IPARTS] = incl;
        if (c0) s_gbase[2 * threadIdx.x] = part_off[2 * threadIdx.x] + atomicAdd(&part_fill[2 * threadIdx.x], (unsigned long long)c0);
        if (c1) s_gbase[2 * threadIdx.x + 1] = part_off[2 * threadIdx.x + 1] + atomicAdd(&part_fill[2 * threadIdx.x + 1], (unsigned long long)c1);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PI; i++)
        if (pp[i] != 0xffu)
        {
            const u32 o = s_start[pp[i]] + rk[i];
            sX[o] = x[i]; sRec[o] = rc[i]; sP[o] = (u8)pp[i];
        }
    __syncthreads();
    const u32 total = s_start[IPARTS];
    for (u32 i = threadIdx.x; i < total; i += PT)
    {
        const u32 q = sP[i];
        const u64 d = s_gbase[q] + (i - s_start[q]);
        outX[d] = sX[i];
        outRec[d] = sRec[i];
    }
}

// records of one partition -> their buckets, in arrival order, into a staging copy of hs (tmp) together with their X.
// Every thread carries 4 records: the chain load X -> load dir -> atomic -> store is pure latency, the 4 chains overlap.
__global__ void __launch_bounds__(256) k_idx_place(const u32 * __restrict__ X, const u64 * __restrict__ rec, u64 n, const i32 * __restrict__ dir,
                                                   u32 * __restrict__ fill, u64 * __restrict__ tmp, u32 * __restrict__ tmpx)
{
    const u64 i0 = (u64)blockIdx.x * (blockDim.x * 4) + threadIdx.x;
    u32 x[4]; i32 b[4], e[4]; u64 r[4];
#pragma unroll
    for (int q = 0; q < 4; q++)
    {
        const u64 i = i0 + (u64)q * blockDim.x;
        x[q] = i < n ? __ldg(X + i) : 0xffffffffu;
        r[q] = i < n ? __ldg(rec + i) : 0;
    }
#pragma unroll
    for (int q = 0; q < 4; q++) { b[q] = 0; e[q] = 0; if (x[q] != 0xffffffffu) { b[q] = __ldg(dir + x[q]); e[q] = __ldg(dir + x[q] + 1); } }
    u32 slot[4];
#pragma unroll
    for (int q = 0; q < 4; q++) slot[q] = e[q] > b[q] ? atomicAdd(&fill[x[q]], 1u) : 0u;
#pragma unroll
    for (int q = 0; q < 4; q++)
        if (e[q] > b[q])
        {
            const u64 o = (u64)b[q] + slot[q];
            tmp[o] = r[q];
            tmpx[o] = x[q];
        }
}
// ascending order inside each bucket (index_util.cpp:1788-1796) by rank: one thread per staged record counts the records
// of its bucket that are smaller (records are distinct) and stores itself at that rank of the final hs, its Y byte next to
// it. Neighbouring threads mostly share a bucket, so the bucket is read once per warp, from L2 (the partition was staged a
// moment ago); the work per bucket is n^2 compares, but uniform and parallel -- a thread-per-bucket insertion sort spent
// 10 ms of the build diverging.
__global__ void __launch_bounds__(256) k_idx_rank(const u64 * __restrict__ tmp, const u32 * __restrict__ tmpx, u64 p0, u64 n, const i32 * __restrict__ dir,
                                                  u64 * __restrict__ hs, u8 * __restrict__ hsy)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 o = p0 + i;
    const u64 r = tmp[o];
    const u32 x = tmpx[o];
    const i32 b = __ldg(dir + x), e = __ldg(dir + x + 1);
    u32 rank = 0;
    for (i32 k = b; k < e; k++) rank += tmp[k] < r ? 1u : 0u;
    hs[(u64)b + rank] = r;
    if (hsy) hsy[(u64)b + rank] = (u8)(r & 0xff);
}

// One thread per sample; a warp owns 32 consecutive samples. A sample whose X differs from the previous sample's X
// (xpre rule, :1882) fetches its bucket's lookup sector and scans the Y keys. The matches of the warp are written, in
// sample order, as one contiguous run of 32-bit entries (hs index | query strand << 31) claimed with ONE atomic per warp
// from one of kMaskPools pools; the fill pass then spreads those entries evenly over the lanes -- every lane keeps a
// random hs load in flight -- instead of walking one sample's matches per thread. A warp that gets no room (pool
// exhausted) is re-scanned by the fill pass. Per sample only the match count is stored (it feeds the device scan).
#ifndef LNR_COUNT_MIN_CTAS
#define LNR_COUNT_MIN_CTAS 8     // 32 registers (24 bytes spilled), 64 warps per SM: 4.65 -> 4.46 ms per 65 536 reads (call 35)
#endif
__global__ void __launch_bounds__(256, LNR_COUNT_MIN_CTAS) k_seed_count(const u8 * __restrict__ bases, const u64 * __restrict__ read_off,
                                                    const SeedTask * __restrict__ tasks, const SeedTaskInfo * __restrict__ tinfo, u32 n_tasks,
                                                    u64 n_samples, const uint4 * __restrict__ dirx, const u8 * __restrict__ hsy,
                                                    u32 * __restrict__ count, SeedWarpRec * wrec,
                                                    u32 * __restrict__ list, u32 pool_cap, unsigned int * pool_ctr)
{
    u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    u32 c = 0, scanned = 0;
    i32 bkt_b = 0; u32 qY = 0;
    uint4 e0 = make_uint4(0, 0, 0, 0), e1 = make_uint4(0, 0, 0, 0);   // the bucket's lookup entry (k_idx_dirx)
    bool active = false;
    u32 X = 0xffffffffu, ti = 0xffffffffu, m = 0;
    SeedTask t;
    memset(&t, 0, sizeof t);
    GAcc acc = {nullptr, 0};
    SeedVal sv = {0, 0, 0};
    // the 32 samples of a warp are consecutive and a task has ~1300: one search per warp, the lanes walk on from it
    const u32 t_first = s - lane < n_samples ? wrec[(s - lane) >> 5].task : 0u;     // laid out by k_seed_prep
    if (s < n_samples)
    {
        // the warp's first task, its read and the start of the next task in one round trip; only the lanes of a warp that
        // straddles a task boundary (1 warp in ~40) walk on
        ti = t_first;
        const u64 nxt = ti + 1 < n_tasks ? tasks[ti + 1].sample0 : ~0ULL;
        t = tasks[ti];
        SeedTaskInfo inf = tinfo[ti];
        if (nxt <= s)
        {
            do ti++; while (ti + 1 < n_tasks && tasks[ti + 1].sample0 <= s);
            t = tasks[ti]; inf = tinfo[ti];
        }
        m = (u32)(s - t.sample0) + 1;
        acc.s = bases + inf.base; acc.len = (i64)inf.len;
        u32 k;
        if (!seed_sample_fast(acc.s, acc.len, t, m, sv)) seed_sample(acc, t, m, sv, k);
        X = sv.X;
    }
    // X of the previous sample of the same task: the lane below usually holds it, otherwise it is recomputed
    u32 xl = __shfl_up_sync(0xffffffffu, X, 1), til = __shfl_up_sync(0xffffffffu, ti, 1), ml = __shfl_up_sync(0xffffffffu, m, 1);
    if (s < n_samples)
    {
        u32 xprev = 0;
        if (m > 1)
        {
            if (lane > 0 && til == ti && ml == m - 1) xprev = xl;
            else { SeedVal pv; u32 kp; if (!seed_sample_fast(acc.s, acc.len, t, m - 1, pv)) seed_sample(acc, t, m - 1, pv, kp); xprev = pv.X; }
        }
        if (sv.X != xprev)
        {
            e0 = ld_dirx(dirx + (size_t)kDirxQuads * sv.X); e1 = ld_dirx(dirx + (size_t)kDirxQuads * sv.X + 1);
            bkt_b = (i32)e0.x;
            qY = sv.Y;
            scanned = e0.y;
            active = scanned != 0;
        }
    }
    // 4 Y bytes per step; ykey_match(b, Y) with v = b ^ Y, l = lowest set bit of v: v == 0 or v >> ctz(v) < 4  <=>  v <= 3 l
    const u32 Y4 = qY * 0x01010101u;
    // all four bytes at once: v <= 3 l  <=>  v has no two set bits two or more positions apart. s = every position at least
    // two below a set bit of v (per byte), d = v & s is zero exactly for the matching bytes; d <= 0x3f, so d + 0x7f sets a
    // byte's top bit iff the byte is non-zero and never carries into the next one; one multiply gathers the four flags
    auto match4raw = [&](u32 w4) -> u32 {
        const u32 v4 = w4 ^ Y4;
        u32 sm = (v4 >> 2) & 0x3f3f3f3fu;
        sm |= (sm >> 1) & 0x7f7f7f7fu;
        sm |= (sm >> 2) & 0x3f3f3f3fu;
        sm |= (sm >> 4) & 0x0f0f0f0fu;
        const u32 d = v4 & sm;
        const u32 z = (((d + 0x7f7f7f7fu) & 0x80808080u) ^ 0x80808080u) >> 7;
        return (z * 0x01020408u) >> 24;
    };
    auto match4 = [&](u32 w4, u32 i) -> u32 {       // the same with the keys behind the bucket's end masked off
        u32 h4 = match4raw(w4);
        u32 rem = scanned - i;
        if (rem < 4) h4 &= (1u << rem) - 1u;
        return h4;
    };
    u64 m0 = 0;   // matches among the bucket's first 64 records; longer buckets (rare) are scanned again when emitting
    const u8 * pb = hsy + bkt_b + kDirxKeys;
    const unsigned sh = ((unsigned)(uintptr_t)pb & 3u) * 8u;
    const u32 * pw = (const u32 *)((uintptr_t)pb & ~(uintptr_t)3);
    if (active)
    {
        // records 0..23 come with the first half of the lookup entry, 24..55 with its second half (the same 64-byte DRAM
        // burst: an L2 hit), only longer buckets go on in hsy
        const u32 ew[6] = {e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
        // the first 64 records: unmasked flags into m0, the bucket's end is applied once at the end (one popcount)
#pragma unroll
        for (int q = 0; q < 6; q++)
            if (4u * q < scanned) m0 |= (u64)match4raw(ew[q]) << (4 * q);
        if (scanned > 24)
        {
            const uint4 e2 = ld_dirx(dirx + (size_t)kDirxQuads * sv.X + 2), e3 = ld_dirx(dirx + (size_t)kDirxQuads * sv.X + 3);
            const u32 fw[8] = {e2.x, e2.y, e2.z, e2.w, e3.x, e3.y, e3.z, e3.w};
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (24u + 4u * q < scanned) m0 |= (u64)match4raw(fw[q]) << (24 + 4 * q);
        }
        if (scanned > (u32)kDirxKeys)
        {
            u32 w0 = __ldg(pw);
            for (u32 i = kDirxKeys; i < scanned; i += 4)
            {
                u32 w1 = __ldg(pw + ((i - kDirxKeys) >> 2) + 1);
                if (i < 64) m0 |= (u64)match4raw(__funnelshift_r(w0, w1, sh)) << i;
                else c += __popc(match4(__funnelshift_r(w0, w1, sh), i));
                w0 = w1;
            }
        }
        if (scanned < 64) m0 &= (1ULL << scanned) - 1ULL;
        c += __popcll(m0);
    }
    // claim the warp's run of entries: exclusive prefix of the per-lane counts, one atomic per warp
    u32 incl = c, hsum = scanned;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 t2 = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += t2; }
#pragma unroll
    for (int o = 16; o; o >>= 1) hsum += __shfl_xor_sync(0xffffffffu, hsum, o);
    const u32 wtot = __shfl_sync(0xffffffffu, incl, 31);
    const unsigned wid = threadIdx.x >> 5;
    const u32 pool = (blockIdx.x * 8u + wid) & (kMaskPools - 1);
    u32 wbase = 0;
    if (lane == 0 && wtot) wbase = atomicAdd(pool_ctr + pool, wtot);
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    const bool have = (u64)wbase + wtot <= (u64)pool_cap;
    if (lane == 0 && s < n_samples)
    {
        SeedWarpRec r; r.list_off = wtot == 0 ? 0u : (have ? pool * pool_cap + wbase : 0xffffffffu); r.scanned = hsum; r.task = t_first; r.pad = 0;
        wrec[s >> 5] = r;
    }
    if (s < n_samples) count[s] = c;
    if (c && have)
    {
        u32 * o = list + (size_t)pool * pool_cap + wbase + (incl - c);
        const u32 tag = (u32)bkt_b | (sv.strand << 31);
        while (m0) { int bit = __ffsll((long long)m0) - 1; m0 &= m0 - 1; *o++ = tag + (u32)bit; }
        if (scanned > 64)
        {
            u32 w0 = __ldg(pw + 2);
            for (u32 i = 64; i < scanned; i += 4)
            {
                u32 w1 = __ldg(pw + ((i - kDirxKeys) >> 2) + 1);
                u32 h4 = match4(__funnelshift_r(w0, w1, sh), i);
                while (h4) { int bit = __ffs((int)h4) - 1; h4 &= h4 - 1; *o++ = tag + i + (u32)bit; }
                w0 = w1;
            }
        }
    }
}
// H (bucket records scanned) of a seeding pass for lnr_last_batch_counters: summed from the per-warp records by a pass of
// its own, so that the count kernel neither synchronises its CTA nor funnels atomics into one line
__global__ void __launch_bounds__(256) k_seed_stats(const SeedWarpRec * __restrict__ wrec, u64 n_warps, unsigned long long * counters)
{
    unsigned long long hs = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_warps; i += (u64)gridDim.x * blockDim.x) hs += wrec[i].scanned;
    for (int o = 16; o; o >>= 1) hs += __shfl_xor_sync(0xffffffffu, hs, o);
    __shared__ unsigned long long s_h[8];
    if ((threadIdx.x & 31) == 0) s_h[threadIdx.x >> 5] = hs;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int i = 1; i < 8; i++) hs += s_h[i];
        if (hs) atomicAdd(&counters[1], hs);
    }
}
// anchors of task ti start at aoff[sample0] + ti (slot 0 of the region is the sentinel). A warp owns the same 32 samples
// as in the count pass; its matches are dealt round-robin to the lanes: entry j of the warp's run belongs to the sample
// whose exclusive count prefix is the largest one <= j (found with 5 shuffles), that sample's k / L / task come by shuffle.
#ifndef LNR_FILL_MIN_CTAS
#define LNR_FILL_MIN_CTAS 8      // 32 registers, 64 warps per SM: the kernel lives on loads in flight, 3.16 -> 2.91 ms (call 35)
#endif
__global__ void __launch_bounds__(256, LNR_FILL_MIN_CTAS) k_seed_fill(const u8 * __restrict__ bases, const u64 * __restrict__ read_off,
                                                   const SeedTask * __restrict__ tasks, const SeedTaskInfo * __restrict__ tinfo, u32 n_tasks,
                                                   u64 n_samples, const u64 * __restrict__ hs, const uint4 * __restrict__ dirx,
                                                   const u64 * __restrict__ aoff, u64 * __restrict__ anchors,
                                                   const u32 * __restrict__ count, const SeedWarpRec * __restrict__ wrec,
                                                   const u32 * __restrict__ list, unsigned long long * counters)
{
    const u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    const u64 s0 = s - lane;
    if (s0 >= n_samples) return;
    const u32 c = s < n_samples ? count[s] : 0u;
    u32 incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 t2 = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += t2; }
    const u32 T = __shfl_sync(0xffffffffu, incl, 31);
    if (T == 0) return;
    const u32 e = incl - c;
    const SeedWarpRec wr = wrec[s0 >> 5];        // (the task of the warp's first sample is here)
    const u64 a0 = aoff[s0];
    // the first 64 entries of the warp's run are on their way before the task of the sample is looked up
    const u32 * lst = list + (wr.list_off != 0xffffffffu ? wr.list_off : 0u);
    u32 pe0 = 0, pe1 = 0;
    if (wr.list_off != 0xffffffffu)
    {
        if (lane < T) pe0 = __ldg(lst + lane);
        if (lane + 32 < T) pe1 = __ldg(lst + lane + 32);
    }
    u32 ti = wr.task, k = 0, L = 0;
    SeedTask t;
    memset(&t, 0, sizeof t);
    if (c)
    {
        const u64 nxt = ti + 1 < n_tasks ? tasks[ti + 1].sample0 : ~0ULL;
        t = tasks[ti];
        L = (u32)tinfo[ti].len;
        if (nxt <= s)
        {
            do ti++; while (ti + 1 < n_tasks && tasks[ti + 1].sample0 <= s);
            t = tasks[ti]; L = (u32)tinfo[ti].len;
        }
        const u32 m = (u32)(s - t.sample0) + 1;
        k = t.str + kSpanD + t.alpha * m - 1;
    }
    if (wr.list_off != 0xffffffffu)
    {
        // two entries per lane and iteration (a warp has ~60): both entry loads, then both random record loads, are in flight
        // together -- the kernel waits on DRAM round trips (long scoreboard 23 cycles per issue), not on anything it computes
        for (u32 base = 0; base < T; base += 64)
        {
            const u32 j0 = base + lane, j1 = j0 + 32;
            const bool two = base + 32 < T;                      // warp-uniform
            const u32 ent0 = base == 0 ? pe0 : (j0 < T ? __ldg(lst + j0) : 0u);
            const u32 ent1 = base == 0 ? pe1 : (j1 < T ? __ldg(lst + j1) : 0u);
            u32 lo0 = 0, lo1 = 0;
#pragma unroll
            for (int step = 16; step; step >>= 1)
            {
                const u32 c0 = lo0 + step;
                const u32 ec0 = __shfl_sync(0xffffffffu, e, c0 & 31);
                if (ec0 <= j0) lo0 = c0;         // e is non-decreasing over the lanes and cand < 32
            }
            if (two)
            {
#pragma unroll
                for (int step = 16; step; step >>= 1)
                {
                    const u32 c1 = lo1 + step;
                    const u32 ec1 = __shfl_sync(0xffffffffu, e, c1 & 31);
                    if (ec1 <= j1) lo1 = c1;
                }
            }
            const u64 h0 = j0 < T ? ld_hs(hs + (ent0 & 0x7fffffffu)) : 0ull;
            const u64 h1 = j1 < T ? ld_hs(hs + (ent1 & 0x7fffffffu)) : 0ull;
            const u32 ko0 = __shfl_sync(0xffffffffu, k, lo0), Lo0 = __shfl_sync(0xffffffffu, L, lo0), tio0 = __shfl_sync(0xffffffffu, ti, lo0);
            if (j0 < T) anchors[a0 + j0 + tio0 + 1] = val2anchor(h0, (u64)ko0, (u64)Lo0, ent0 >> 31);
            if (two)
            {
                const u32 ko1 = __shfl_sync(0xffffffffu, k, lo1), Lo1 = __shfl_sync(0xffffffffu, L, lo1), tio1 = __shfl_sync(0xffffffffu, ti, lo1);
                if (j1 < T) anchors[a0 + j1 + tio1 + 1] = val2anchor(h1, (u64)ko1, (u64)Lo1, ent1 >> 31);
            }
        }
    }
    else if (c)
    {
        // the warp got no room for its entries in the count pass: evaluate the sample again and scan its bucket in hs
        atomicAdd(&counters[18], 1ULL);   // diagnostics: samples re-scanned
        GAcc acc = {bases + tinfo[ti].base, (i64)L};
        SeedVal sv; u32 kk;
        const u32 m = (u32)(s - t.sample0) + 1;
        if (!seed_sample_fast(acc.s, acc.len, t, m, sv)) seed_sample(acc, t, m, sv, kk);
        const uint4 d0 = __ldg(dirx + (size_t)kDirxQuads * sv.X);
        u64 * out = anchors + a0 + e + ti + 1;
        for (u32 i = 0; i < d0.y; i++)
        {
            u64 h = __ldg(hs + d0.x + i);
            if (ykey_match((u32)(h & kMaskY), sv.Y)) *out++ = val2anchor(h, (u64)k, (u64)L, sv.strand);
        }
    }
}

#include "lnr_hindex.cuh"

// =====================================================================================================
// warp-per-read pipeline kernels
// =====================================================================================================
struct ReadSlot
{
    u32 n_cords;     // cords so far / final
    u32 status;      // 0 done, 1 waiting for re-map, 2 capacity failure
    u32 task0, n_tasks;   // re-map tasks
};
struct MapArgs
{
    const u64 * read_off; const u8 * bases; u32 n_reads;
    const F96 * feats; const u64 * foff;               // read features (feature_type 1: shorts behind the cast)
    const F96 * const * f2; const u32 * nf2;           // genome features
    int ft; u32 win;                                   // feature type (2 = 2_48, 1 = 1_32) and its window size
    const SeedTask * tasks; u32 n_tasks;               // seeding tasks of this pass
    const u64 * aoff;                                  // anchor offsets per sample (+ n_samples sentinel entry)
    u64 * A; u64 * B;
    u64 * cords; const u64 * cords_base;               // per-read cord regions
    ReadSlot * slots;
    SeedTask * tasks2; u32 tasks2_cap; u32 * n_tasks2; // re-map tasks produced by the primary pass
    u32 * bins; u8 * arena; u64 arena_per_warp;
    u64 fit_cap;                                       // k_hits_sort: the smallest per-warp arena of the three section kernels
    // Heavy tasks of the primary pass (more raw anchors than fit_n: their scratch bound exceeds a warp's share of the arena).
    // k_order_tasks lists them; in each of the three section kernels the first kBigWarps warps take them first, each on its
    // slot of the big arena, so the few reads with the most anchors start at once instead of queueing behind everything else
    u32 fit_n; const u32 * heavy_list; const u32 * n_heavy; u32 * queue_h;
    u32 * queue;                                       // atomic work counter
    const u32 * order;                                 // reads sorted by length, longest first (tail latency)
    int index_type;                                    // 1 DIndex, 2 HIndex (sample grid of re-map tasks)
    u32 * task_nhits;                                  // hits per task after stage 1 (0xffffffff = scratch exhausted)
    u32 * task_state;                                  // primary pass, between the hit sections: n2 after k_hits_sort, hits after
                                                       // k_hits_chain; 0 = the task is finished, 0xffffffff = left to the big-arena pass
    u32 * big_list; u32 * n_big;                       // tasks whose scratch did not fit the per-warp arena: re-run with the big arena
    u8 * big_arena; u64 big_arena_per_warp;
    u64 * warp_rec;                                    // optional: 16 u64 per warp, profile of its slowest task
    u64 warp_rec_stage_off;                            // u64 offset of the section kernels' records (8 per warp and section)
    u64 warp_rec_stage_stride;                         // warps per section in that area
    float stop_ratio;
    unsigned long long * counters;
    // optional debug
    u64 * dbg_hits; const u64 * dbg_hoff; u32 * dbg_nhits; u64 * dbg_c1; u32 * dbg_nc1;
};

__device__ __forceinline__ void fill_pipe_in(const MapArgs & a, u32 r, PipeIn & in)
{
    u64 L = a.read_off[r + 1] - a.read_off[r];
    in.read = a.bases + a.read_off[r];
    in.L = (u32)L;
    in.ft = a.ft; in.win = a.win;
    in.f1[0] = in.f1[1] = nullptr; in.s1[0] = in.s1[1] = nullptr;
    if (a.ft == 1)
    {
        in.nf1 = feat32_count(L);
        in.s1[0] = (const i16 *)a.feats + a.foff[r];
        in.s1[1] = in.s1[0] + in.nf1;
    }
    else
    {
        in.nf1 = feat_count_read(L);
        in.f1[0] = a.feats + a.foff[r];
        in.f1[1] = in.f1[0] + in.nf1;
    }
    in.f2 = a.f2; in.s2 = (const i16 * const *)a.f2; in.nf2 = a.nf2;
    in.stop_ratio = a.stop_ratio;
}

// scratch of the cord-block chaining stage for a read with n cords
struct FinishBufs { Blk * sp1; Blk * sp2; i32 * sc1; i32 * sc2; BlockScratch s1, s2; u64 * tmp; };
__device__ __forceinline__ bool alloc_finish(Arena & ar, FinishBufs & f, int n)
{
    int cap = n + 2;
    f.sp1 = arena_alloc<Blk>(ar, cap);
    f.sp2 = arena_alloc<Blk>(ar, cap);
    f.sc1 = arena_alloc<i32>(ar, cap);
    f.sc2 = arena_alloc<i32>(ar, cap);
    block_scratch_alloc(ar, f.s1, cap);
    block_scratch_alloc(ar, f.s2, cap);
    f.tmp = arena_alloc<u64>(ar, (u64)cap);
    return !ar.failed;
}
// finish a read (all lanes): gather blocks again if it was re-mapped, then chainBlocksCords + flags
__device__ __noinline__ void finish_read(const Warp & w, FinishBufs & f, u64 L, u64 * cords, int & nc, Blk * sep_in, int n_sep_in, bool regather, u32 win)
{
    int n_sep;
    if (regather)
    {
        int dummy = 0;
        n_sep = gather_blocks_w(w, cords, nc, (YPair *)0, dummy, f.sp1, L, 1000, win, 1);
    }
    else
    {
        n_sep = n_sep_in;
        for (int i = w.lane; i < n_sep; i += 32) f.sp1[i] = sep_in[i];
        __syncwarp();
    }
    phase_finish_w(w, L, cords, nc, f.sp1, n_sep, f.sp2, f.sc1, f.sc2, f.s1, f.s2, f.tmp, win);
    nc = __shfl_sync(0xffffffffu, nc, 0);
    __syncwarp();
}

// Processing order of the primary pass: tasks with the most raw anchors first (their hits stage is the longest, so
// they must not start last). Counting sort into 132 quarter-octave buckets, one CTA.
__global__ void __launch_bounds__(1024) k_order_tasks(const SeedTask * __restrict__ tasks, u32 n_tasks, const u64 * __restrict__ aoff,
                                                      u32 * __restrict__ order, u32 fit_n = 0xffffffffu, u32 * heavy_list = nullptr,
                                                      u32 * n_heavy = nullptr)
{
    __shared__ u32 s_cnt[136];
    for (int i = threadIdx.x; i < 136; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    auto bucket = [&](u32 t) -> int {
        u64 n = aoff[tasks[t].sample0 + tasks[t].n_samples] - aoff[tasks[t].sample0];
        if (n == 0) return 0;
        int lg = 63 - __clzll((long long)n);
        int frac = lg >= 2 ? (int)((n >> (lg - 2)) & 3) : 0;
        int b = lg * 4 + frac + 1;
        return b > 135 ? 135 : b;
    };
    for (u32 t = threadIdx.x; t < n_tasks; t += blockDim.x)
    {
        atomicAdd(&s_cnt[bucket(t)], 1u);
        if (heavy_list)     // n of phase_map = raw anchors + the sentinel slot
        {
            u64 n = aoff[tasks[t].sample0 + tasks[t].n_samples] - aoff[tasks[t].sample0] + 1;
            if (n > (u64)fit_n) heavy_list[atomicAdd(n_heavy, 1u)] = t;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        u32 run = 0;
        for (int b = 135; b >= 0; b--) { u32 c = s_cnt[b]; s_cnt[b] = run; run += c; }
    }
    __syncthreads();
    for (u32 t = threadIdx.x; t < n_tasks; t += blockDim.x) order[atomicAdd(&s_cnt[bucket(t)], 1u)] = t;
}

// the same ordering by an explicit per-task key (the hit-section kernels: anchors after the filters, chained hits)
__global__ void __launch_bounds__(1024) k_order_by_key(const u32 * __restrict__ key, u32 n_tasks, u32 * __restrict__ order)
{
    __shared__ u32 s_cnt[136];
    for (int i = threadIdx.x; i < 136; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    auto bucket = [&](u32 t) -> int {
        u32 n = key[t];
        if (n == 0 || n == 0xffffffffu) return 0;
        int lg = 31 - __clz((int)n);
        int frac = lg >= 2 ? (int)((n >> (lg - 2)) & 3) : 0;
        return lg * 4 + frac + 1;
    };
    for (u32 t = threadIdx.x; t < n_tasks; t += blockDim.x) atomicAdd(&s_cnt[bucket(t)], 1u);
    __syncthreads();
    if (threadIdx.x == 0)
    {
        u32 run = 0;
        for (int b = 135; b >= 0; b--) { u32 c = s_cnt[b]; s_cnt[b] = run; run += c; }
    }
    __syncthreads();
    for (u32 t = threadIdx.x; t < n_tasks; t += blockDim.x) order[atomicAdd(&s_cnt[bucket(t)], 1u)] = t;
}

// ---- stage 1: hits. One warp per seeding task (primary pass: task r = read r; re-map pass: one task per gap).
// Everything of apxMap_ up to and including _filterHits; the hits replace the task's anchors in A.
#ifndef LNR_HITS_MIN_CTAS
#define LNR_HITS_MIN_CTAS 6
#endif
__global__ void __launch_bounds__(128, LNR_HITS_MIN_CTAS) k_map_hits(MapArgs a, int remap_pass, int big_pass)
{
    // big_pass: second, tiny launch over the tasks whose scratch bound exceeds the per-warp arena (reads with very many
    // anchors), a few warps with a large arena each
    __shared__ u32 s_hist[4][kWarpSmemWords];
    for (int i = threadIdx.x; i < 4 * kWarpSmemWords; i += blockDim.x) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    Warp w = {(int)(threadIdx.x & 31), 0xffffffffu};
    u32 wid = threadIdx.x >> 5;
    u32 gw = blockIdx.x * (blockDim.x >> 5) + wid;
    Arena ar = {a.arena + (u64)gw * a.arena_per_warp, a.arena_per_warp, 0, 0};
    if (big_pass) { ar.base = a.big_arena + (u64)gw * a.big_arena_per_warp; ar.cap = a.big_arena_per_warp; }
    const u32 n_units = big_pass ? *a.n_big : a.n_tasks;
    u32 * bins = a.bins + (u64)gw * kNumBins;
    PipeCounters cnt;
    memset(&cnt, 0, sizeof cnt);
    const bool want_rec = a.warp_rec && !remap_pass && !big_pass;
    u64 g_start = 0, g_task = 0, n_done = 0, g_busy = 0;
    if (want_rec) g_start = globaltimer_ns();
    while (true)
    {
        u32 q = 0;
        if (w.lane == 0) q = atomicAdd(a.queue, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= n_units) break;
        if (want_rec) { g_task = globaltimer_ns(); n_done++; }
        u32 ti = big_pass ? a.big_list[q] : (remap_pass ? q : a.order[q]);   // primary: heaviest reads first
        if (big_pass && w.lane == 0) atomicAdd(&a.counters[16], 1ULL);       // diagnostics: tasks taken by the big-arena pass
        const SeedTask t = a.tasks[ti];
        u32 r = t.read;
        u64 L = a.read_off[r + 1] - a.read_off[r];
        if (w.lane == 0) a.task_nhits[ti] = 0;
        if (L <= (u64)kMinReadLen) continue;            // mapper.cpp:440
        long long t_read = LNR_CLOCK();
        PipeCounters before = cnt;
        PipeIn in;
        fill_pipe_in(a, r, in);
        u64 s0 = t.sample0, s1 = s0 + t.n_samples;
        u64 base = a.aoff[s0] + ti;
        int n = (int)(a.aoff[s1] - a.aoff[s0]) + 1;
        int nc_dummy = 0;
        u64 * dh = (!remap_pass && a.dbg_hits) ? a.dbg_hits + a.dbg_hoff[r] : (u64 *)0;
        u32 dcap = dh ? (u32)(a.dbg_hoff[r + 1] - a.dbg_hoff[r]) : 0;
        int rc = phase_map(w, ar, s_hist[wid], bins, in, a.A + base, a.B + base, n, (u64)t.str, remap_pass ? ((u64)t.end & kMaskY) : (L & kMaskY),
                           remap_pass ? 1 : 0, (u64 *)0, nc_dummy, 0, dh, dh ? a.dbg_nhits + r : (u32 *)0, dcap, cnt, a.A + base,
                           a.task_nhits + ti, big_pass != 0);
        if (rc && w.lane == 0)
        {
            a.task_nhits[ti] = 0xffffffffu;                 // unresolved (big pass pending) or failed for good
            if (rc == 2) a.big_list[atomicAdd(a.n_big, 1u)] = ti;   // untouched: re-run with the big arena
        }
        if (!remap_pass) cnt.t[12]++;
        u64 dt = (u64)(LNR_CLOCK() - t_read);
        cnt.t[14] += dt;
        if (want_rec) g_busy += globaltimer_ns() - g_task;
        if (dt > cnt.t[13])
        {
            cnt.t[13] = dt;
            if (a.warp_rec && !remap_pass && !big_pass && w.lane == 0)   // this warp's slowest task so far (tail analysis)
            {
                u64 * rec = a.warp_rec + (u64)gw * 24;
                for (int i = 0; i < 12; i++) rec[i] = cnt.t[i] - before.t[i];
                rec[12] = dt; rec[13] = r; rec[14] = L; rec[15] = (u64)n; rec[18] = g_task; rec[20] = q;
            }
        }
    }
    if (want_rec && w.lane == 0)
    {
        u64 * rec = a.warp_rec + (u64)gw * 24;
        rec[16] = g_start; rec[17] = globaltimer_ns(); rec[19] = n_done; rec[21] = g_busy; rec[22] = cnt.t[14];
    }
    if (w.lane == 0)
    {
        if (cnt.hits) atomicAdd(&a.counters[3], (unsigned long long)cnt.hits);
        for (int i = 0; i < 16; i++)
            if (cnt.t[i])
            {
                if (i == 13) atomicMax(&a.counters[24 + i], (unsigned long long)cnt.t[i]);
                else atomicAdd(&a.counters[24 + i], (unsigned long long)cnt.t[i]);
            }
    }
}

// ---- primary pass of stage 1 as three kernels, one per section of the hit stage (lnr_pipeline.h hits_sec_*). Per-task
// state between them lives in the task's own anchor regions: k_hits_sort leaves the x-sorted anchors in A[0..n2),
// k_hits_chain the chained hits in B[0..n_hits) and their scores (int32) in A, k_hits_blocks the final hits in A.
static const u32 kBigWarps = 32;      // slots of the big arena
struct StageCommon
{
    Warp w; u32 gw; Arena ar; PipeCounters cnt;
    Arena small, big; bool big_warp, heavy_open;
    // LNR_LONGEST_PROFILE: when this warp started / retired and its slowest task
    u64 * rec; u64 g_start; long long t_task; u64 max_dur, max_ti, max_size, max_q, q, n_done;
};
__device__ __forceinline__ void stage_begin(const MapArgs & a, StageCommon & c)
{
    c.w = {(int)(threadIdx.x & 31), 0xffffffffu};
    c.gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    c.ar = {a.arena + (u64)c.gw * a.arena_per_warp, a.arena_per_warp, 0, 0};
    c.small = c.ar;
    c.big_warp = a.heavy_list != nullptr && c.gw < kBigWarps;
    c.heavy_open = c.big_warp;
    c.big = {a.big_arena + (u64)c.gw * a.big_arena_per_warp, a.big_arena_per_warp, 0, 0};
    memset(&c.cnt, 0, sizeof c.cnt);
    c.rec = nullptr;
}
__device__ __forceinline__ void stage_rec_begin(const MapArgs & a, StageCommon & c, int stage)
{
    if (!a.warp_rec || (u64)c.gw >= a.warp_rec_stage_stride) return;     // (a section launched with more warps than the record area has rows)
    c.rec = a.warp_rec + ((u64)stage * a.warp_rec_stage_stride + c.gw) * 8 + (u64)a.warp_rec_stage_off;
    c.g_start = globaltimer_ns(); c.max_dur = 0; c.max_ti = 0; c.max_size = 0; c.max_q = 0; c.n_done = 0; c.t_task = 0;
}
__device__ __forceinline__ void stage_task_done(StageCommon & c, u32 ti, u64 size)
{
    if (!c.rec) return;
    u64 d = (u64)(clock64() - c.t_task);
    c.n_done++;
    if (d > c.max_dur) { c.max_dur = d; c.max_ti = ti; c.max_size = size; c.max_q = c.q; }
}
// next task of this warp; heavy = it comes from the heavy list and runs on the warp's big-arena slot (c.ar is switched).
// Heavy tasks that a warp meets in the ordinary queue are skipped by the caller (task_is_heavy): the big warps own them.
__device__ __forceinline__ bool stage_next(const MapArgs & a, StageCommon & c, u32 & ti, bool & heavy)
{
    heavy = false;
    if (c.heavy_open)
    {
        u32 h = 0;
        if (c.w.lane == 0) h = atomicAdd(a.queue_h, 1u);
        h = __shfl_sync(0xffffffffu, h, 0);
        if (h < *a.n_heavy)
        {
            ti = a.heavy_list[h];
            heavy = true;
            c.ar = c.big;
            if (c.rec) { c.q = h; c.t_task = clock64(); }
            return true;
        }
        c.heavy_open = false;
        c.ar = c.small;
    }
    u32 q = 0;
    if (c.w.lane == 0) q = atomicAdd(a.queue, 1u);
    q = __shfl_sync(0xffffffffu, q, 0);
    if (q >= a.n_tasks) return false;
    ti = a.order[q];                     // heaviest reads first
    if (c.rec) { c.q = q; c.t_task = clock64(); }
    return true;
}
__device__ __forceinline__ bool task_is_heavy(const MapArgs & a, int n) { return a.heavy_list != nullptr && (u32)n > a.fit_n; }
__device__ __forceinline__ void stage_end(const MapArgs & a, const StageCommon & c)
{
    if (c.rec && c.w.lane == 0)
    {
        c.rec[0] = c.g_start; c.rec[1] = globaltimer_ns(); c.rec[2] = c.max_dur; c.rec[3] = c.max_ti; c.rec[4] = c.max_size;
        c.rec[5] = c.max_q; c.rec[6] = c.n_done;
    }
    if (c.w.lane == 0)
    {
        if (c.cnt.hits) atomicAdd(&a.counters[3], (unsigned long long)c.cnt.hits);
        for (int i = 0; i < 16; i++)
            if (c.cnt.t[i])
            {
                if (i == 13) atomicMax(&a.counters[24 + i], (unsigned long long)c.cnt.t[i]);
                else atomicAdd(&a.counters[24 + i], (unsigned long long)c.cnt.t[i]);
            }
    }
}
__device__ __forceinline__ void task_region(const MapArgs & a, u32 ti, const SeedTask & t, u64 & base, int & n)
{
    u64 s0 = t.sample0, s1 = s0 + t.n_samples;
    base = a.aoff[s0] + ti;
    n = (int)(a.aoff[s1] - a.aoff[s0]) + 1;
}

__global__ void __launch_bounds__(128, 8) k_hits_sort(MapArgs a)
{
    __shared__ u32 s_hist[4][kWarpSmemWords];
    for (int i = threadIdx.x; i < 4 * kWarpSmemWords; i += blockDim.x) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    StageCommon c;
    stage_begin(a, c);
    const Warp w = c.w;
    u32 * bins = a.bins + (u64)c.gw * kNumBins;
    stage_rec_begin(a, c, 0);
    u32 ti; bool heavy;
    while (stage_next(a, c, ti, heavy))
    {
        const SeedTask t = a.tasks[ti];
        u32 r = t.read;
        u64 L = a.read_off[r + 1] - a.read_off[r];
        u64 base; int n;
        task_region(a, ti, t, base, n);
        if (!heavy && task_is_heavy(a, n)) continue;    // taken from the heavy list by one of the big warps
        if (heavy && w.lane == 0) atomicAdd(&a.counters[19], 1ULL);    // diagnostics: tasks served by the heavy lane
        if (w.lane == 0) { a.task_nhits[ti] = 0; a.task_state[ti] = 0; }
        if (L <= (u64)kMinReadLen) continue;            // mapper.cpp:440
        long long tl = LNR_CLOCK();
        if (w.lane == 0 && a.dbg_hits)
        {
            a.dbg_nhits[r] = 1;
            if (a.dbg_hoff[r + 1] > a.dbg_hoff[r]) a.dbg_hits[a.dbg_hoff[r]] = kFlagEnd;
        }
        c.cnt.t[12]++;
        if (heavy ? phase_map_scratch_bound(n) > c.ar.cap : (phase_map_scratch_bound(n) > a.fit_cap || (u64)n * 16 + 4096 > c.ar.cap))
        {
            // does not fit the per-warp arena: the whole task is left, untouched, to the big-arena pass
            if (w.lane == 0) { a.task_nhits[ti] = 0xffffffffu; a.task_state[ti] = 0xffffffffu; a.big_list[atomicAdd(a.n_big, 1u)] = ti; }
            continue;
        }
        arena_reset(c.ar);
        u64 * A = a.A + base;
        u64 * X; int n2;
        int rc = hits_sec_sort(w, c.ar, s_hist[threadIdx.x >> 5], bins, A, a.B + base, n, c.cnt, tl, X, n2);
        if (rc == 0)
        {
            if (X != A)
            {
                for (int i = w.lane; i < n2; i += 32) A[i] = X[i];
            }
            if (w.lane == 0) a.task_state[ti] = (u32)n2;
        }
        else if (rc == 1 && w.lane == 0) { a.task_nhits[ti] = 0xffffffffu; a.task_state[ti] = 0; }
        __syncwarp();
        stage_task_done(c, ti, (u64)n);
    }
    stage_end(a, c);
}


__global__ void __launch_bounds__(128, LNR_CHAIN_MIN_CTAS) k_hits_chain(MapArgs a)
{
    StageCommon c;
    stage_begin(a, c);
    const Warp w = c.w;
    stage_rec_begin(a, c, 1);
    u32 ti; bool heavy;
    while (stage_next(a, c, ti, heavy))
    {
        const SeedTask t = a.tasks[ti];
        u64 base; int n;
        task_region(a, ti, t, base, n);
        if (!heavy && task_is_heavy(a, n)) continue;
        const u32 n2 = a.task_state[ti];
        if (n2 == 0 || n2 == 0xffffffffu) continue;
        long long tl = LNR_CLOCK();
        PipeIn in;
        fill_pipe_in(a, t.read, in);
        arena_reset(c.ar);
        i32 * score = arena_alloc<i32>(c.ar, (u64)n2 + 2);
        int n_hits = 1;
        int rc = c.ar.failed ? 1 : hits_sec_chain(w, c.ar, in, a.A + base, (int)n2, 0, a.B + base, score, n_hits, c.cnt, tl);
        if (rc == 0)
        {
            i32 * dst = (i32 *)(a.A + base);             // the anchors are consumed: A now carries the hit scores
            for (int i = w.lane; i < n_hits; i += 32) dst[i] = score[i];
            if (w.lane == 0) a.task_state[ti] = (u32)n_hits;
        }
        else if (w.lane == 0) { a.task_nhits[ti] = 0xffffffffu; a.task_state[ti] = 0; }
        __syncwarp();
        stage_task_done(c, ti, (u64)n2);
    }
    stage_end(a, c);
}

__global__ void __launch_bounds__(128, LNR_BLOCKS_MIN_CTAS) k_hits_blocks(MapArgs a)
{
    __shared__ u32 s_hist[4][256];
    StageCommon c;
    stage_begin(a, c);
    const Warp w = c.w;
    stage_rec_begin(a, c, 2);
    u32 ti; bool heavy;
    while (stage_next(a, c, ti, heavy))
    {
        const SeedTask t = a.tasks[ti];
        u64 base; int n;
        task_region(a, ti, t, base, n);
        if (!heavy && task_is_heavy(a, n)) continue;
        const u32 nh = a.task_state[ti];
        if (nh == 0 || nh == 0xffffffffu) continue;
        const u32 r = t.read;
        long long tl = LNR_CLOCK();
        PipeIn in;
        fill_pipe_in(a, r, in);
        arena_reset(c.ar);
        int n_hits = (int)nh;
        i32 * score = arena_alloc<i32>(c.ar, (u64)n_hits + 2);
        int rc = 1;
        if (!c.ar.failed)
        {
            const i32 * src = (const i32 *)(a.A + base);
            for (int i = w.lane; i < n_hits; i += 32) score[i] = src[i];
            __syncwarp();
            u64 * dh = a.dbg_hits ? a.dbg_hits + a.dbg_hoff[r] : (u64 *)0;
            u32 dcap = dh ? (u32)(a.dbg_hoff[r + 1] - a.dbg_hoff[r]) : 0;
            rc = hits_sec_blocks(w, c.ar, s_hist[threadIdx.x >> 5], in, a.B + base, score, n_hits, a.A + base, dh, dh ? a.dbg_nhits + r : (u32 *)0, dcap, c.cnt, tl);
            if (rc == 0 && w.lane == 0) a.task_nhits[ti] = (u32)n_hits;
        }
        if (rc == 1 && w.lane == 0) a.task_nhits[ti] = 0xffffffffu;
        __syncwarp();
        stage_task_done(c, ti, (u64)nh);
    }
    stage_end(a, c);
}

// ---- stage 2: window extension (path_dst_2 + extendWindow), one warp per read on a regular grid; reads are taken in
// size order so that neighbouring warps have similar trip counts.
// 12 CTAs (48 warps) per SM: the walk is a serial chain of window steps, more resident warps hide it better than more
// registers do (64 regs / 8 CTAs: 5.10 ms per 65 536 reads; 40 regs / 12 CTAs: 4.82; 32 regs / 16 CTAs: 5.01 -- call 28)
#ifndef LNR_EXTEND_MIN_CTAS
#define LNR_EXTEND_MIN_CTAS 12
#endif
#define LNR_EXTEND_BOUNDS __launch_bounds__(128, LNR_EXTEND_MIN_CTAS)
__global__ void LNR_EXTEND_BOUNDS k_map_extend(MapArgs a, const u32 * __restrict__ read_list, u32 n_list, int remap_pass, int group)
{
    // the lanes of a warp cooperate on one read: they share the 18 script distances of a window step (sub-warp groups
    // and one thread per read were measured slower; `group` is kept in the signature and must be 32)
    u32 tid = blockIdx.x * blockDim.x + threadIdx.x;
    u32 q = tid / 32u;
    int gl = (int)(tid & 31u);
    (void)group;
    PipeCounters cnt;
    memset(&cnt, 0, sizeof cnt);
    long long t0 = LNR_CLOCK();
    if (q < n_list)
    {
        u32 r = read_list[q];
        u64 L = a.read_off[r + 1] - a.read_off[r];
        if (L > (u64)kMinReadLen)
        {
            Warp w1 = {gl, 0xffffffffu};
            PipeIn in;
            fill_pipe_in(a, r, in);
            ReadSlot slot = a.slots[r];
            u64 * cords = a.cords + a.cords_base[r];
            int cap = (int)(a.cords_base[r + 1] - a.cords_base[r]);
            int nc = remap_pass ? (int)slot.n_cords : 0;
            u32 t_first = remap_pass ? slot.task0 : r, t_cnt = remap_pass ? slot.n_tasks : 1;
            bool ok = remap_pass ? slot.status == 1 : true;
            for (u32 g = 0; g < t_cnt && ok; g++)
            {
                u32 ti = t_first + g;
                const SeedTask t = a.tasks[ti];
                u32 nh = a.task_nhits[ti];
                if (nh == 0xffffffffu) { ok = false; break; }
                u64 base = a.aoff[t.sample0] + ti;
                const u64 rend = remap_pass ? ((u64)t.end & kMaskY) : (L & kMaskY);
                ok = a.ft == 1 ? path_dst_2<1>(w1, in, a.A + base, (int)nh, cords, nc, cap, (u64)t.str, rend, cnt)
                               : path_dst_2<2>(w1, in, a.A + base, (int)nh, cords, nc, cap, (u64)t.str, rend, cnt);
            }
            if (gl == 0)
            {
                slot.n_cords = ok ? (u32)nc : 0;
                if (!ok) slot.status = 2;
                a.slots[r] = slot;
            }
        }
    }
    cnt.t[8] = (u64)(LNR_CLOCK() - t0);
    u64 wn = gl == 0 ? cnt.windows : 0, tc = cnt.t[8];
    for (int o = 16; o; o >>= 1) { wn += __shfl_xor_sync(0xffffffffu, wn, o); tc = max(tc, __shfl_xor_sync(0xffffffffu, tc, o)); }
    if ((threadIdx.x & 31) == 0)
    {
        if (wn) atomicAdd(&a.counters[4], (unsigned long long)wn);
        atomicAdd(&a.counters[24 + 8], (unsigned long long)tc);
    }
}

// ---- stage 3: clean / gaps / re-map decision / cord-block chaining, one warp per read
__global__ void __launch_bounds__(128) k_map_finish(MapArgs a, const u32 * __restrict__ read_list, u32 n_list, int remap_pass, int big_pass)
{
    // big_pass: tiny second launch over the reads whose chaining scratch did not fit the per-warp arena
    Warp w = {(int)(threadIdx.x & 31), 0xffffffffu};
    u32 wid = threadIdx.x >> 5;
    u32 gw = blockIdx.x * (blockDim.x >> 5) + wid;
    Arena ar = {a.arena + (u64)gw * a.arena_per_warp, a.arena_per_warp, 0, 0};
    if (big_pass) { ar.base = a.big_arena + (u64)gw * a.big_arena_per_warp; ar.cap = a.big_arena_per_warp; }
    const u32 n_units = big_pass ? *a.n_big : n_list;
    PipeCounters cnt;
    memset(&cnt, 0, sizeof cnt);
    u64 c_cords = 0;
    while (true)
    {
        u32 q = 0;
        if (w.lane == 0) q = atomicAdd(a.queue, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= n_units) break;
        u32 r = big_pass ? a.big_list[q] : read_list[q];
        if (big_pass && w.lane == 0) atomicAdd(&a.counters[17], 1ULL);       // diagnostics: reads finished by the big-arena pass
        u64 L = a.read_off[r + 1] - a.read_off[r];
        ReadSlot slot = a.slots[r];
        if (L <= (u64)kMinReadLen) { slot.n_cords = 0; slot.status = 0; slot.task0 = 0; slot.n_tasks = 0; if (w.lane == 0) a.slots[r] = slot; continue; }
        if (slot.status == 2) continue;
        u64 * cords = a.cords + a.cords_base[r];
        int nc = (int)slot.n_cords;
        int rc = 0;
        long long tl = LNR_CLOCK();
        // all scratch of this stage is claimed before anything is modified, so that a read that does not fit can be
        // handed to the big-arena pass untouched
        arena_reset(ar);
        YPair * str_ends = arena_alloc<YPair>(ar, (u64)nc + 2);
        Blk * sep = arena_alloc<Blk>(ar, (u64)nc + 2);
        int gcap = (int)(L / 1000 + 4);
        YPair * gaps = arena_alloc<YPair>(ar, (u64)gcap);
        FinishBufs fb;
        alloc_finish(ar, fb, nc);
        if (ar.failed)
        {
            if (w.lane == 0)
            {
                if (!big_pass) a.big_list[atomicAdd(a.n_big, 1u)] = r;
                else { slot.n_cords = 0; slot.status = 2; a.slots[r] = slot; }
            }
            continue;
        }
        if (!remap_pass)
        {
            if (a.dbg_c1)
            {
                for (int i = w.lane; i < nc; i += 32) a.dbg_c1[a.cords_base[r] + i] = cords[i];
                if (w.lane == 0) a.dbg_nc1[r] = (u32)nc;
            }
            int remap = 0, n_sep = 0, n_gaps = 0;
            u32 task0 = 0;
            remap = phase_mid_w(w, L, cords, nc, str_ends, sep, n_sep, gaps, n_gaps, gcap, a.win);
            if (w.lane == 0 && remap == 1)
            {
                task0 = atomicAdd(a.n_tasks2, (u32)n_gaps);
                if (task0 + (u32)n_gaps > a.tasks2_cap) remap = -1;
                else
                    for (int i = 0; i < n_gaps; i++)
                    {
                        SeedTask t2;
                        t2.read = r; t2.str = (u32)(gaps[i].first & kMaskY); t2.end = (u32)gaps[i].second; t2.alpha = 7;
                        t2.n_samples = a.index_type == 2 ? hseed_task_samples(t2.str, t2.end, 7) : seed_task_samples(t2.str, t2.end, 7);
                        t2.bias = 0; t2.kskip = 0; t2.pad = 0; t2.sample0 = 0;
                        a.tasks2[task0 + i] = t2;
                    }
            }
            remap = __shfl_sync(0xffffffffu, remap, 0);
            task0 = __shfl_sync(0xffffffffu, task0, 0);
            __syncwarp();
            if (remap < 0) rc = 1;
            LNR_LAP(cnt, 9, tl);
            if (!rc && remap == 0) finish_read(w, fb, L, cords, nc, sep, n_sep, false, a.win);
            LNR_LAP(cnt, 10, tl);
            slot.n_cords = rc ? 0 : (u32)nc;
            slot.status = rc ? 2u : (remap == 1 ? 1u : 0u);
            slot.task0 = task0; slot.n_tasks = remap == 1 ? (u32)n_gaps : 0;
            if (!rc && remap == 0) c_cords += (u64)nc;
        }
        else
        {
            finish_read(w, fb, L, cords, nc, (Blk *)0, 0, true, a.win);
            LNR_LAP(cnt, 10, tl);
            slot.n_cords = (u32)nc;
            slot.status = 0u;
            c_cords += (u64)nc;
        }
        if (w.lane == 0) a.slots[r] = slot;
    }
    if (w.lane == 0)
    {
        if (c_cords) atomicAdd(&a.counters[5], (unsigned long long)c_cords);
        for (int i = 0; i < 16; i++) if (cnt.t[i]) atomicAdd(&a.counters[24 + i], (unsigned long long)cnt.t[i]);
    }
}

// ---- -c 0 (apxMap with f_chain = 0, alg_type 1): one warp per read runs a whole attempt -- sort, run list, hits, path_dst_1
// (phase_c0). attempt 0 = the primary tasks (k-mer step 15); a read whose longest block stays below 0.7 of its length
// (c0_attempt_too_short) becomes a task of the second attempt (step 7, GetDHitListParms 20 / 1), which the host seeds and
// runs through this kernel again; whichever attempt is the read's last also cleans the blocks and sets the flags.
__global__ void __launch_bounds__(128) k_map_c0(MapArgs a, int attempt, int big_pass, int list_n, int best_n)
{
    __shared__ u32 s_hist[4][kWarpSmemWords];
    for (int i = threadIdx.x; i < 4 * kWarpSmemWords; i += blockDim.x) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    Warp w = {(int)(threadIdx.x & 31), 0xffffffffu};
    u32 wid = threadIdx.x >> 5;
    u32 gw = blockIdx.x * (blockDim.x >> 5) + wid;
    Arena ar = {a.arena + (u64)gw * a.arena_per_warp, a.arena_per_warp, 0, 0};
    if (big_pass) { ar.base = a.big_arena + (u64)gw * a.big_arena_per_warp; ar.cap = a.big_arena_per_warp; }
    const u32 n_units = big_pass ? *a.n_big : a.n_tasks;
    PipeCounters cnt;
    memset(&cnt, 0, sizeof cnt);
    u64 c_cords = 0;
    while (true)
    {
        u32 q = 0;
        if (w.lane == 0) q = atomicAdd(a.queue, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= n_units) break;
        u32 ti = big_pass ? a.big_list[q] : (attempt ? q : a.order[q]);
        if (big_pass && w.lane == 0) atomicAdd(&a.counters[16], 1ULL);
        const SeedTask t = a.tasks[ti];
        const u32 r = t.read;
        const u64 L = a.read_off[r + 1] - a.read_off[r];
        ReadSlot slot;
        slot.n_cords = 0; slot.status = 0; slot.task0 = 0; slot.n_tasks = 0;
        if (L <= (u64)kMinReadLen) { if (w.lane == 0) a.slots[r] = slot; continue; }   // mapper.cpp:440
        PipeIn in;
        fill_pipe_in(a, r, in);
        u64 base; int n;
        task_region(a, ti, t, base, n);
        u64 * cords = a.cords + a.cords_base[r];
        const int cap = (int)(a.cords_base[r + 1] - a.cords_base[r]);
        int nc = 0;
        u64 max_len = 0;
        int rc = a.ft == 1 ? phase_c0<1>(w, ar, s_hist[wid], in, a.A + base, a.B + base, n, list_n, best_n, cords, nc, cap, cnt, max_len, big_pass != 0)
                           : phase_c0<2>(w, ar, s_hist[wid], in, a.A + base, a.B + base, n, list_n, best_n, cords, nc, cap, cnt, max_len, big_pass != 0);
        if (rc == 2)
        {
            if (w.lane == 0) a.big_list[atomicAdd(a.n_big, 1u)] = ti;      // untouched: re-run with the big arena
            continue;
        }
        if (rc) { slot.status = 2; if (w.lane == 0) a.slots[r] = slot; continue; }
        if (attempt == 0 && c0_attempt_too_short(max_len, L, a.win))
        {
            // clear(cords_str); toggle(1); second attempt (pmpfinder.cpp:2781-2784)
            u32 t0 = 0;
            if (w.lane == 0)
            {
                t0 = atomicAdd(a.n_tasks2, 1u);
                if (t0 < a.tasks2_cap)
                {
                    SeedTask t2;
                    t2.read = r; t2.str = 0; t2.end = (u32)L; t2.alpha = 7;
                    t2.n_samples = a.index_type == 2 ? hseed_task_samples(0, (u32)L, 7) : seed_task_samples(0, (u32)L, 7);
                    t2.bias = 0; t2.kskip = 0; t2.pad = 0; t2.sample0 = 0;
                    a.tasks2[t0] = t2;
                }
            }
            t0 = __shfl_sync(0xffffffffu, t0, 0);
            slot.status = t0 < a.tasks2_cap ? 1u : 2u;
            slot.task0 = t0; slot.n_tasks = 1;
            if (w.lane == 0) a.slots[r] = slot;
            continue;
        }
        c0_finish(w, L, cords, nc, a.win);
        slot.n_cords = (u32)nc;
        c_cords += (u64)nc;
        if (w.lane == 0) a.slots[r] = slot;
    }
    if (w.lane == 0)
    {
        if (cnt.hits) atomicAdd(&a.counters[3], (unsigned long long)cnt.hits);
        if (cnt.windows) atomicAdd(&a.counters[4], (unsigned long long)cnt.windows);
        if (c_cords) atomicAdd(&a.counters[5], (unsigned long long)c_cords);
    }
}

__global__ void k_slot_counts(const ReadSlot * __restrict__ slots, u32 n, u32 * __restrict__ cnt, u32 * __restrict__ n_fail)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    cnt[i] = slots[i].n_cords;
    if (slots[i].status == 2) atomicAdd(n_fail, 1u);
}
// one warp per read copies its cords to the concatenated output
__global__ void k_gather_cords(const u64 * __restrict__ cords, const u64 * __restrict__ cords_base, const u32 * __restrict__ ncords,
                               const u64 * __restrict__ out_off, u32 n_reads, u64 * __restrict__ out, u64 out_cap)
{
    u32 r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    u32 lane = threadIdx.x & 31;
    if (r >= n_reads) return;
    u64 o = out_off[r];
    const u64 * src = cords + cords_base[r];
    u32 n = ncords[r];
    for (u32 i = lane; i < n; i += 32)
        if (o + i < out_cap) out[o + i] = src[i];
}

// ---- scan helper ------------------------------------------------------------------------------------------
// out[i] = exclusive prefix of in[0..n) (after optional capping); *d_total (device u64) receives the sum
template <class OutT>
static int device_scan(lnr_ctx * ctx, u32 * d_in, u64 n, u32 cap, OutT * d_out, u64 * d_total, const char * tag)
{
    u32 nb = (u32)((n + STILE - 1) / STILE);
    CK(ctx->scan_tmp.reserve(((size_t)nb + 2) * sizeof(u64)));
    u64 * part = ctx->scan_tmp.as<u64>();
    {
        LaunchScope ls(ctx, tag, 3);
        k_scan_reduce<<<nb, ST, 0, ctx->stream>>>(d_in, n, cap, part);
        k_scan_spine<<<1, 1024, 0, ctx->stream>>>(part, nb, d_total);
        k_scan_apply<OutT><<<nb, ST, 0, ctx->stream>>>(d_in, n, part, d_out);
    }
    CK(cudaGetLastError());
    return LNR_OK;
}


// =====================================================================================================
// C ABI
// =====================================================================================================
struct SelfKeyHi { __device__ u64 operator()(u64 v) const { return v >> 32; } };
__global__ void k_selftest_sort(u64 * a, u64 * s0, u64 * s1, int n, u64 * out)
{
    __shared__ u32 hist[256];
    Warp w = {(int)threadIdx.x, 0xffffffffu};
    u64 * r = gnu_sort_w(w, hist, a, s0, s1, n, 32, SelfKeyHi());
    for (int i = threadIdx.x; i < n; i += 32) out[i] = r[i];
}

// ---- SAM* / BAM* records from cords (lnr_bamrec.h): one thread per read, count -> scan -> fill -------------------------
template <bool FILL>
__global__ void __launch_bounds__(128) k_bam_records(const u64 * __restrict__ cords, const u64 * __restrict__ cords_off, const u64 * __restrict__ read_len,
                                                     u32 n_reads, BamParms P, u32 * __restrict__ n_rec, u32 * __restrict__ n_cig,
                                                     const u64 * __restrict__ rec_off, const u64 * __restrict__ cig_off, BamRec * __restrict__ recs,
                                                     u64 * __restrict__ cigs)
{
    const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const u64 * cs = cords + cords_off[r];
    const u32 n = (u32)(cords_off[r + 1] - cords_off[r]);
    if (!FILL)
    {
        BamCountSink sink;
        bam_walk(cs, n, read_len[r], P, sink);
        n_rec[r] = sink.n_rec; n_cig[r] = sink.n_cig;
    }
    else
    {
        BamFillSink sink;
        sink.recs = recs + rec_off[r]; sink.cig = cigs + cig_off[r];
        bam_walk(cs, n, read_len[r], P, sink);
    }
}

#include "lnr_ingest.cuh"

extern "C" {

int lnr_ctx_create(int device, lnr_ctx ** out)
{
    if (!out) return LNR_E_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return LNR_E_CUDA;
    lnr_ctx * ctx = new lnr_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return LNR_E_CUDA; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    ctx->n_sm = prop.multiProcessorCount;
    if (const char * e = getenv("LNR_MAP_CTAS_PER_SM")) { int v = atoi(e); if (v >= 1 && v <= 16) ctx->map_ctas_per_sm = v; }
    if (const char * e = getenv("LNR_SORT_CTAS_PER_SM")) { int v = atoi(e); if (v >= 1 && v <= 8) ctx->sort_ctas_per_sm = v; }
    if (const char * e = getenv("LNR_CHAIN_CTAS_PER_SM")) { int v = atoi(e); if (v >= 1 && v <= 12) ctx->chain_ctas_per_sm = v; }
    if (const char * e = getenv("LNR_BLOCKS_CTAS_PER_SM")) { int v = atoi(e); if (v >= 1 && v <= 12) ctx->blocks_ctas_per_sm = v; }
    if (const char * e = getenv("LNR_BIG_ARENA_MB")) { int v = atoi(e); if (v >= 1 && v <= 16384) ctx->big_arena_bytes_per_warp = (size_t)v << 20; }
    if (const char * e = getenv("LNR_ARENA_KB")) { int v = atoi(e); if (v >= 16 && v <= (1 << 20)) ctx->arena_bytes_per_warp = (size_t)v << 10; }
    else if (const char * e = getenv("LNR_ARENA_MB")) { int v = atoi(e); if (v >= 1 && v <= 1024) ctx->arena_bytes_per_warp = (size_t)v << 20; }
    if (const char * e = getenv("LNR_L2_FETCH")) { int v = atoi(e); if (v == 32 || v == 64 || v == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)v); }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return LNR_E_CUDA; }
    {
        if (getenv("LNR_TRACE"))
        {
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_map_hits, 128, 0);
            fprintf(stderr, "[lnr trace] occupancy API: k_map_hits %d CTAs/SM", nb);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_map_finish, 128, 0);
            fprintf(stderr, ", k_map_finish %d", nb);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_map_extend, 128, 0);
            fprintf(stderr, ", k_map_extend %d", nb);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_feat_reads, FT, 0);
            fprintf(stderr, ", k_feat_reads %d\n", nb);
        }
    }
    *out = ctx;
    return LNR_OK;
}
void lnr_ctx_destroy(lnr_ctx * ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    harvest_stats(ctx);
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    for (DevBuf * b : {&ctx->bases, &ctx->read_off, &ctx->tasks, &ctx->sample_info, &ctx->sample_cnt, &ctx->scan_tmp, &ctx->anchorsA,
                       &ctx->anchorsB, &ctx->feats, &ctx->foff, &ctx->ftile, &ctx->cords, &ctx->cords_base, &ctx->ncords, &ctx->slots,
                       &ctx->bins, &ctx->arena, &ctx->tasks2, &ctx->misc, &ctx->out_cords, &ctx->out_off, &ctx->dbg_hits, &ctx->dbg_hoff,
                       &ctx->dbg_nhits, &ctx->dbg_c1, &ctx->dbg_nc1, &ctx->read_meta, &ctx->remap_list, &ctx->ing[0], &ctx->ing[1], &ctx->ing[2], &ctx->ing[3], &ctx->ing[4], &ctx->ing[5], &ctx->ing[6], &ctx->ing[7], &ctx->order, &ctx->order2, &ctx->mask_ctr, &ctx->tile_read, &ctx->task_nhits, &ctx->task_state, &ctx->big_arena, &ctx->big_list, &ctx->seed_masks, &ctx->seed_mask_off, &ctx->warp_rec, &ctx->packed, &ctx->task_info, &ctx->feat_pairs})
        b->release();
    ctx->stage.release();
    if (ctx->reads_cache) cudaFree(ctx->reads_cache);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}
const char * lnr_last_error(const lnr_ctx * ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int lnr_ctx_set_profiling(lnr_ctx * ctx, int on) { if (!ctx) return LNR_E_ARG; ctx->profiling = on != 0; return LNR_OK; }
int lnr_ctx_reset_kernel_times(lnr_ctx * ctx)
{
    if (!ctx) return LNR_E_ARG;
    cudaStreamSynchronize(ctx->stream);
    harvest_stats(ctx);
    ctx->stats.clear();
    ctx->stat_order.clear();
    return LNR_OK;
}
int lnr_ctx_kernel_times(lnr_ctx * ctx, int cap, const char ** names, float * total_ms, uint64_t * launches)
{
    if (!ctx) return LNR_E_ARG;
    cudaStreamSynchronize(ctx->stream);
    harvest_stats(ctx);
    int i = 0;
    for (auto & nm : ctx->stat_order)
    {
        if (i < cap)
        {
            auto it = ctx->stats.find(nm);
            names[i] = it->first.c_str();
            total_ms[i] = (float)it->second.ms;
            launches[i] = it->second.launches;
        }
        i++;
    }
    return i;
}

// ---- genome -------------------------------------------------------------------------------------------
static int genome_layout(lnr_ctx * ctx, uint32_t n_contigs, const uint64_t * len, lnr_genome * g)
{
    if (n_contigs == 0 || n_contigs >= 1024) return fail(ctx, LNR_E_LIMIT, "1..1023 contigs (linear.cpp:107)");
    uint64_t off = 0;
    for (uint32_t i = 0; i < n_contigs; i++)
    {
        if (len[i] >= (1ULL << 30) - (1ULL << 20)) return fail(ctx, LNR_E_LIMIT, "contig length must be < 2^30 - 2^20 (cords.cpp:13-15)");
        g->len.push_back(len[i]);
        g->off.push_back(off);
        off += (len[i] + 256 + 255) & ~255ULL;
    }
    g->total_padded = off + 256;
    return LNR_OK;
}
int lnr_genome_upload(lnr_ctx * ctx, uint32_t n_contigs, const uint8_t * const * dna5, const uint64_t * len, lnr_genome ** out)
{
    if (!ctx || !dna5 || !len || !out) return LNR_E_ARG;
    cudaSetDevice(ctx->device);
    lnr_genome * g = new lnr_genome();
    g->ctx = ctx; g->n_contigs = n_contigs; g->d_bases = nullptr;
    int rc = genome_layout(ctx, n_contigs, len, g);
    if (rc) { delete g; return rc; }
    cudaError_t e = cudaMalloc(&g->d_bases, g->total_padded);
    if (e != cudaSuccess) { delete g; return fail(ctx, LNR_E_CUDA, cudaGetErrorString(e)); }
    cudaMemsetAsync(g->d_bases, 0, g->total_padded, ctx->stream);
    for (uint32_t i = 0; i < n_contigs; i++)
        cudaMemcpyAsync(g->d_bases + g->off[i], dna5[i], len[i], cudaMemcpyHostToDevice, ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cudaFree(g->d_bases); delete g; return fail(ctx, LNR_E_CUDA, cudaGetErrorString(e)); }
    *out = g;
    return LNR_OK;
}
int lnr_genome_from_device(lnr_ctx * ctx, uint32_t n_contigs, const uint8_t * dev_concat, const uint64_t * len, lnr_genome ** out)
{
    if (!ctx || !dev_concat || !len || !out) return LNR_E_ARG;
    cudaSetDevice(ctx->device);
    lnr_genome * g = new lnr_genome();
    g->ctx = ctx; g->n_contigs = n_contigs; g->d_bases = nullptr;
    int rc = genome_layout(ctx, n_contigs, len, g);
    if (rc) { delete g; return rc; }
    cudaError_t e = cudaMalloc(&g->d_bases, g->total_padded);
    if (e != cudaSuccess) { delete g; return fail(ctx, LNR_E_CUDA, cudaGetErrorString(e)); }
    cudaMemsetAsync(g->d_bases, 0, g->total_padded, ctx->stream);
    uint64_t src = 0;
    for (uint32_t i = 0; i < n_contigs; i++)
    {
        cudaMemcpyAsync(g->d_bases + g->off[i], dev_concat + src, len[i], cudaMemcpyDeviceToDevice, ctx->stream);
        src += len[i];
    }
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cudaFree(g->d_bases); delete g; return fail(ctx, LNR_E_CUDA, cudaGetErrorString(e)); }
    *out = g;
    return LNR_OK;
}
void lnr_genome_destroy(lnr_genome * g)
{
    if (!g) return;
    cudaSetDevice(g->ctx->device);
    if (g->d_bases) cudaFree(g->d_bases);
    delete g;
}

// ---- genome features ------------------------------------------------------------------------------------
int lnr_features_build(lnr_ctx * ctx, const lnr_genome * g, int feature_type, unsigned threads_sem, lnr_feats ** out)
{
    if (!ctx || !g || !out || threads_sem == 0) return LNR_E_ARG;
    if (feature_type != 2 && feature_type != 1) return fail(ctx, LNR_E_UNSUPPORTED, "feature_type must be 2 (2-mer/48, -f 2) or 1 (1-mer/32, -f 1)");
    cudaSetDevice(ctx->device);
    const size_t esz = feature_type == 1 ? sizeof(i16) : sizeof(F96);
    lnr_feats * f = new lnr_feats();
    f->ctx = ctx; f->feature_type = feature_type; f->n_contigs = g->n_contigs; f->d_f = nullptr; f->d_ptrs = nullptr; f->d_n = nullptr;
    uint64_t tot = 0;
    for (uint32_t i = 0; i < g->n_contigs; i++)
    {
        uint32_t n = feature_type == 1 ? feat32_count(g->len[i]) : feat_count_genome(g->len[i], threads_sem);
        f->n.push_back(n);
        f->off.push_back(tot);
        tot += (n + 8 + 7) & ~7ull;   // a little slack between contigs; strings stay 16-byte aligned for either entry size
    }
    CK(cudaMalloc(&f->d_f, (tot + 8) * esz));
    CK(cudaMemsetAsync(f->d_f, 0, (tot + 8) * esz, ctx->stream));
    CK(cudaMalloc(&f->d_ptrs, g->n_contigs * sizeof(F96 *)));
    CK(cudaMalloc(&f->d_n, g->n_contigs * sizeof(u32)));
    std::vector<const F96 *> ptrs;
    for (uint32_t i = 0; i < g->n_contigs; i++) ptrs.push_back((const F96 *)((const u8 *)f->d_f + f->off[i] * esz));
    CK(cudaMemcpyAsync(f->d_ptrs, ptrs.data(), ptrs.size() * sizeof(F96 *), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(f->d_n, f->n.data(), f->n.size() * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
    for (uint32_t i = 0; i < g->n_contigs; i++)
    {
        if (!f->n[i]) continue;
        LaunchScope ls(ctx, "k_feat_genome");
        if (feature_type == 1)
        {
            const u32 nw = feat32_written_parallel(g->len[i]);   // the entries behind stay 0 (canonical rule)
            if (nw) k_feat32_genome<<<(nw + (FT - 1) - 1) / (FT - 1), FT, 0, ctx->stream>>>(g->d_bases + g->off[i], (i64)g->len[i], nw, (i16 *)f->d_f + f->off[i]);
            continue;
        }
        u32 grid = (f->n[i] + FE - 1) / FE;
        k_feat_genome<<<grid, FT, 0, ctx->stream>>>(g->d_bases + g->off[i], (i64)g->len[i], f->n[i], f->d_f + f->off[i]);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    *out = f;
    return LNR_OK;
}
int lnr_features_count(const lnr_feats * f, uint32_t contig, uint64_t * n)
{
    if (!f || !n || contig >= f->n_contigs) return LNR_E_ARG;
    *n = f->n[contig];
    return LNR_OK;
}
int lnr_features_download(const lnr_feats * f, uint32_t contig, void * dst, uint64_t cap_entries, uint64_t * n_entries)
{
    if (!f || contig >= f->n_contigs) return LNR_E_ARG;
    lnr_ctx * ctx = f->ctx;
    cudaSetDevice(ctx->device);
    if (n_entries) *n_entries = f->n[contig];
    if (!dst) return LNR_OK;
    if (cap_entries < f->n[contig]) return fail(ctx, LNR_E_CAPACITY, "feature buffer too small");
    const size_t esz = f->feature_type == 1 ? sizeof(i16) : sizeof(F96);
    CK(cudaMemcpy(dst, (const u8 *)f->d_f + f->off[contig] * esz, (size_t)f->n[contig] * esz, cudaMemcpyDeviceToHost));
    return LNR_OK;
}
void lnr_features_destroy(lnr_feats * f)
{
    if (!f) return;
    cudaSetDevice(f->ctx->device);
    if (f->d_f) cudaFree(f->d_f);
    if (f->d_ptrs) cudaFree((void *)f->d_ptrs);
    if (f->d_n) cudaFree(f->d_n);
    delete f;
}

// ---- multi-GPU index build over NCCL ---------------------------------------------------------------------
// NCCL is bound at run time (dlopen of libnccl.so.2: in a torch.distributed process that is the library already loaded,
// in a plain C++ host the system one), so liblnr_b200.so has no link-time dependency on it.
namespace {
typedef struct ncclComm * nccl_comm_t;
struct NcclUid { char internal[128]; };
struct NcclApi
{
    void * lib = nullptr;
    int (*GetUniqueId)(NcclUid *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, NcclUid, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char * (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi & nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char * name : {"libnccl.so.2", "libnccl.so"})
        {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) return;
#define LNR_NCCL_SYM(field, sym) *(void **)(&api.field) = dlsym(api.lib, sym)
        LNR_NCCL_SYM(GetUniqueId, "ncclGetUniqueId"); LNR_NCCL_SYM(CommInitRank, "ncclCommInitRank"); LNR_NCCL_SYM(CommDestroy, "ncclCommDestroy");
        LNR_NCCL_SYM(AllGather, "ncclAllGather"); LNR_NCCL_SYM(Broadcast, "ncclBroadcast"); LNR_NCCL_SYM(GroupStart, "ncclGroupStart");
        LNR_NCCL_SYM(GroupEnd, "ncclGroupEnd"); LNR_NCCL_SYM(GetErrorString, "ncclGetErrorString");
        LNR_NCCL_SYM(Send, "ncclSend"); LNR_NCCL_SYM(Recv, "ncclRecv");
#undef LNR_NCCL_SYM
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.Broadcast && api.GroupStart && api.GroupEnd && api.Send && api.Recv;
    });
    return api;
}
const int kNcclInt32 = 2, kNcclUint64 = 5;   // ncclDataType_t values (nccl.h)
}  // namespace

struct lnr_comm { lnr_ctx * ctx; nccl_comm_t comm; int rank, n_ranks; bool owned; };

#define CKN(call)                                                                                                  \
    do {                                                                                                           \
        int r_ = (call);                                                                                           \
        if (r_ != 0) {                                                                                             \
            ctx->err = std::string(#call) + ": " + (nccl_api().GetErrorString ? nccl_api().GetErrorString(r_) : "NCCL error"); \
            return LNR_E_CUDA;                                                                                     \
        }                                                                                                          \
    } while (0)

int lnr_nccl_unique_id(uint8_t id[128])
{
    if (!id) return LNR_E_ARG;
    NcclApi & api = nccl_api();
    if (!api.ok) return LNR_E_UNSUPPORTED;
    NcclUid u;
    if (api.GetUniqueId(&u) != 0) return LNR_E_CUDA;
    memcpy(id, u.internal, 128);
    return LNR_OK;
}
int lnr_comm_create(lnr_ctx * ctx, const uint8_t id[128], int rank, int n_ranks, lnr_comm ** out)
{
    if (!ctx || !id || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return LNR_E_ARG;
    NcclApi & api = nccl_api();
    if (!api.ok) return fail(ctx, LNR_E_UNSUPPORTED, "libnccl.so.2 not found");
    cudaSetDevice(ctx->device);
    NcclUid u;
    memcpy(u.internal, id, 128);
    nccl_comm_t c = nullptr;
    CKN(api.CommInitRank(&c, n_ranks, u, rank));
    *out = new lnr_comm{ctx, c, rank, n_ranks, true};
    return LNR_OK;
}
int lnr_comm_from_nccl(lnr_ctx * ctx, void * nccl_comm, int rank, int n_ranks, lnr_comm ** out)
{
    if (!ctx || !nccl_comm || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return LNR_E_ARG;
    if (!nccl_api().ok) return fail(ctx, LNR_E_UNSUPPORTED, "libnccl.so.2 not found");
    *out = new lnr_comm{ctx, (nccl_comm_t)nccl_comm, rank, n_ranks, false};
    return LNR_OK;
}
void lnr_comm_destroy(lnr_comm * c)
{
    if (!c) return;
    if (c->owned && c->comm) { cudaSetDevice(c->ctx->device); nccl_api().CommDestroy(c->comm); }
    delete c;
}

__global__ void k_gather_i32(const i32 * __restrict__ src, const u32 * __restrict__ idx, u32 n, i32 * __restrict__ out)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}
__global__ void k_idx_rebase(i32 * __restrict__ dir, u32 x0, u32 n, i32 base)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dir[x0 + i] += base;
}

struct ShardPlan { lnr_comm * comm; };   // non-null: build one minimizer range and assemble over NCCL
static int dindex_build(lnr_ctx * ctx, const lnr_genome * g, unsigned threads_sem, u32 x_lo, u32 x_hi, lnr_comm * comm, lnr_index ** out);
static int hindex_build(lnr_ctx * ctx, const lnr_genome * g, unsigned T, lnr_comm * comm, lnr_index ** out);

// X ranges of the sharded HIndex build: cut r = the first X at which the running pair count reaches r/n_ranks of all pairs
// (pure host arithmetic on the per-X histogram every rank computes identically; exported so that the CPU tests can check it)
int lnr_hindex_shard_cuts(const uint32_t * pairs_per_x, uint32_t n_x, int n_ranks, uint32_t * cuts)
{
    if (!pairs_per_x || !cuts || n_ranks < 1) return LNR_E_ARG;
    u64 all = 0;
    for (u32 x = 0; x < n_x; x++) all += pairs_per_x[x];
    for (int r = 0; r <= n_ranks; r++) cuts[r] = n_x;
    cuts[0] = 0;
    u64 run = 0; int r = 1;
    for (u32 x = 0; x < n_x && r < n_ranks; x++)
    {
        while (r < n_ranks && run >= (all * (u64)r + (u64)n_ranks - 1) / (u64)n_ranks) cuts[r++] = x;
        run += pairs_per_x[x];
    }
    return LNR_OK;
}

int lnr_index_build_sharded(lnr_ctx * ctx, const lnr_genome * g, int index_type, unsigned threads_sem, lnr_comm * comm, lnr_index ** out)
{
    if (!ctx || !g || !comm || !out || threads_sem == 0) return LNR_E_ARG;
    if (index_type == 2) return hindex_build(ctx, g, threads_sem, comm, out);
    if (index_type != 1) return fail(ctx, LNR_E_UNSUPPORTED, "index_type must be 1 (DIndex, -i 1) or 2 (HIndex, -i 2)");
    if (((1u << kDirBits) % (u32)comm->n_ranks) != 0) return fail(ctx, LNR_E_ARG, "the number of ranks must divide 2^26");
    const u32 per = (1u << kDirBits) / (u32)comm->n_ranks;
    return dindex_build(ctx, g, threads_sem, (u32)comm->rank * per, ((u32)comm->rank + 1) * per, comm, out);
}

// ---- index ---------------------------------------------------------------------------------------------------
static int index_build_range(lnr_ctx * ctx, const lnr_genome * g, int index_type, unsigned threads_sem, u32 x_lo, u32 x_hi, lnr_index ** out);

int lnr_index_build(lnr_ctx * ctx, const lnr_genome * g, int index_type, unsigned threads_sem, lnr_index ** out)
{
    return index_build_range(ctx, g, index_type, threads_sem, 0u, 1u << kDirBits, out);
}
int lnr_index_build_shard(lnr_ctx * ctx, const lnr_genome * g, int index_type, unsigned threads_sem, unsigned shard, unsigned n_shards,
                          lnr_index ** out)
{
    if (n_shards == 0 || shard >= n_shards || ((1u << kDirBits) % n_shards) != 0) return fail(ctx, LNR_E_ARG, "n_shards must divide 2^26");
    u32 per = (1u << kDirBits) / n_shards;
    return index_build_range(ctx, g, index_type, threads_sem, shard * per, (shard + 1) * per, out);
}
int lnr_index_export_dindex_device(const lnr_index * ix, int32_t * dev_dir, uint64_t * dev_hs, uint64_t hs_cap)
{
    if (!ix || ix->index_type != 1) return LNR_E_ARG;
    lnr_ctx * ctx = ix->ctx;
    cudaSetDevice(ctx->device);
    if (dev_dir) CK(cudaMemcpyAsync(dev_dir, ix->d_dir, (size_t)kDirSize * sizeof(i32), cudaMemcpyDeviceToDevice, ctx->stream));
    if (dev_hs)
    {
        if (hs_cap < ix->n_hs) return fail(ctx, LNR_E_CAPACITY, "hs buffer too small");
        CK(cudaMemcpyAsync(dev_hs, ix->d_hs, (size_t)ix->n_hs * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return LNR_OK;
}
static cudaError_t index_build_dirx(lnr_ctx * ctx, lnr_index * ix)
{
    if (ix->d_dirx || ix->index_type != 1) return cudaSuccess;
    cudaError_t e = cudaMalloc(&ix->d_dirx, (size_t)(kDirSize - 1) * 16 * kDirxQuads);
    if (e != cudaSuccess) return e;
    LaunchScope ls(ctx, "k_idx_dirx");
    k_idx_dirx<<<((kDirSize - 1) + 255) / 256, 256, 0, ctx->stream>>>(ix->d_dir, ix->d_hsy, kDirSize - 1, ix->d_dirx);
    return cudaGetLastError();
}

int lnr_index_from_device(lnr_ctx * ctx, const int32_t * dev_dir, const uint64_t * dev_hs, uint64_t n_hs, lnr_index ** out)
{
    if (!ctx || !dev_dir || !out || (n_hs && !dev_hs)) return LNR_E_ARG;
    if (n_hs >= (1ULL << 31)) return fail(ctx, LNR_E_LIMIT, "hs exceeds int32 bucket offsets (index_util.h:101)");
    cudaSetDevice(ctx->device);
    lnr_index * ix = new lnr_index();
    ix->ctx = ctx; ix->index_type = 1; ix->d_dir = nullptr; ix->d_hs = nullptr; ix->n_hs = n_hs;
    if (cudaMalloc(&ix->d_dir, (size_t)kDirSize * sizeof(i32)) != cudaSuccess || cudaMalloc(&ix->d_hs, (size_t)(n_hs + 8) * sizeof(u64)) != cudaSuccess)
    {
        lnr_index_destroy(ix);
        return fail(ctx, LNR_E_CUDA, "cudaMalloc failed in lnr_index_from_device");
    }
    cudaMemcpyAsync(ix->d_dir, dev_dir, (size_t)kDirSize * sizeof(i32), cudaMemcpyDeviceToDevice, ctx->stream);
    if (n_hs) cudaMemcpyAsync(ix->d_hs, dev_hs, (size_t)n_hs * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream);
    if (cudaMalloc(&ix->d_hsy, (size_t)n_hs + 64) != cudaSuccess) { lnr_index_destroy(ix); return fail(ctx, LNR_E_CUDA, "cudaMalloc failed in lnr_index_from_device"); }
    if (n_hs) k_idx_split_y<<<(u32)((n_hs + 255) / 256), 256, 0, ctx->stream>>>(ix->d_hs, n_hs, ix->d_hsy);
    if (index_build_dirx(ctx, ix) != cudaSuccess) { lnr_index_destroy(ix); return fail(ctx, LNR_E_CUDA, "cudaMalloc failed in lnr_index_from_device"); }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { lnr_index_destroy(ix); return fail(ctx, LNR_E_CUDA, cudaGetErrorString(e)); }
    *out = ix;
    return LNR_OK;
}
// ---- index serialisation (include/lnr_b200.h: lnr_index_save / lnr_index_load) ----------------------------------------------
struct IdxFileHeader
{
    char magic[8];               // "LNRIDX1"
    uint32_t index_type, elem_b; // 1 DIndex / 2 HIndex; bytes per element of array B (8 hs record, 16 directory slot)
    uint64_t n_a, n_b;           // elements of array A (dir int32 / ysa u64) and B (hs u64 / directory table)
    uint64_t empty_dir, n_dir_entries;
    uint64_t sum_a, sum_b;       // word-wise checksums
};
static_assert(sizeof(IdxFileHeader) == 64, "index file header is 64 bytes");
static uint64_t idx_file_sum(uint64_t h, const void * p, size_t bytes)
{
    const uint64_t * w = (const uint64_t *)p;
    size_t n = bytes / 8;
    for (size_t i = 0; i < n; i++) h = (h ^ w[i]) * 0x100000001b3ULL;
    const unsigned char * t = (const unsigned char *)p + n * 8;
    for (size_t i = 0; i < bytes - n * 8; i++) h = (h ^ t[i]) * 0x100000001b3ULL;
    return h;
}
static const size_t kIdxFileChunk = (size_t)64 << 20;
// device array -> file (write = true) or file -> device array, 64 MB at a time through a pinned staging buffer
static int idx_file_stream(lnr_ctx * ctx, FILE * f, void * dev, size_t bytes, bool write, void * stage, uint64_t * sum)
{
    uint64_t h = 0xcbf29ce484222325ULL;
    for (size_t o = 0; o < bytes; o += kIdxFileChunk)
    {
        const size_t n = std::min(kIdxFileChunk, bytes - o);
        if (write)
        {
            CK(cudaMemcpyAsync(stage, (const u8 *)dev + o, n, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            if (fwrite(stage, 1, n, f) != n) return fail(ctx, LNR_E_ARG, "index file: write failed");
        }
        else
        {
            if (fread(stage, 1, n, f) != n) return fail(ctx, LNR_E_ARG, "index file: truncated");
            CK(cudaMemcpyAsync((u8 *)dev + o, stage, n, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
        h = idx_file_sum(h, stage, n);
    }
    *sum = h;
    return LNR_OK;
}
int lnr_index_save(const lnr_index * ix, const char * path)
{
    if (!ix || !path) return LNR_E_ARG;
    lnr_ctx * ctx = ix->ctx;
    cudaSetDevice(ctx->device);
    IdxFileHeader hd;
    memset(&hd, 0, sizeof hd);
    memcpy(hd.magic, "LNRIDX1", 8);
    hd.index_type = (uint32_t)ix->index_type;
    void * a; void * b; size_t ba, bb;
    if (ix->index_type == 1)
    {
        hd.n_a = kDirSize; hd.n_b = ix->n_hs; hd.elem_b = 8;
        a = ix->d_dir; ba = (size_t)kDirSize * sizeof(i32); b = ix->d_hs; bb = (size_t)ix->n_hs * sizeof(u64);
    }
    else
    {
        hd.n_a = ix->n_ysa; hd.n_b = ix->tab_len; hd.elem_b = (uint32_t)sizeof(HNode);
        hd.empty_dir = ix->empty_dir; hd.n_dir_entries = ix->n_dir_entries;
        a = ix->d_ysa; ba = (size_t)ix->n_ysa * sizeof(u64); b = ix->d_tab; bb = (size_t)ix->tab_len * sizeof(HNode);
    }
    FILE * f = fopen(path, "wb");
    if (!f) return fail(ctx, LNR_E_ARG, "index file: cannot open for writing");
    void * stage = nullptr;
    if (cudaMallocHost(&stage, kIdxFileChunk) != cudaSuccess) { fclose(f); return fail(ctx, LNR_E_CUDA, "cudaMallocHost failed in lnr_index_save"); }
    int rc = fwrite(&hd, 1, sizeof hd, f) == sizeof hd ? LNR_OK : fail(ctx, LNR_E_ARG, "index file: write failed");
    if (!rc) rc = idx_file_stream(ctx, f, a, ba, true, stage, &hd.sum_a);
    if (!rc) rc = idx_file_stream(ctx, f, b, bb, true, stage, &hd.sum_b);
    if (!rc && (fseek(f, 0, SEEK_SET) != 0 || fwrite(&hd, 1, sizeof hd, f) != sizeof hd)) rc = fail(ctx, LNR_E_ARG, "index file: write failed");
    cudaFreeHost(stage);
    if (fclose(f) != 0 && !rc) rc = fail(ctx, LNR_E_ARG, "index file: write failed");
    return rc;
}
int lnr_index_load(lnr_ctx * ctx, const char * path, lnr_index ** out)
{
    if (!ctx || !path || !out) return LNR_E_ARG;
    cudaSetDevice(ctx->device);
    FILE * f = fopen(path, "rb");
    if (!f) return fail(ctx, LNR_E_ARG, "index file: cannot open");
    IdxFileHeader hd;
    if (fread(&hd, 1, sizeof hd, f) != sizeof hd || memcmp(hd.magic, "LNRIDX1", 8) != 0) { fclose(f); return fail(ctx, LNR_E_ARG, "not an index file (magic LNRIDX1)"); }
    const bool dtype = hd.index_type == 1 && hd.n_a == (uint64_t)kDirSize && hd.elem_b == 8 && hd.n_b < (1ULL << 31);
    const bool htype = hd.index_type == 2 && hd.elem_b == sizeof(HNode) && hd.n_b >= 2 && (hd.n_b & (hd.n_b - 1)) == 0 && hd.n_a >= 2 && hd.n_a < (1ULL << 40);
    if (!dtype && !htype) { fclose(f); return fail(ctx, LNR_E_ARG, "index file: inconsistent header"); }
    lnr_index * ix = new lnr_index();
    ix->ctx = ctx; ix->index_type = (int)hd.index_type; ix->d_dir = nullptr; ix->d_hs = nullptr; ix->n_hs = 0;
    void * stage = nullptr;
    int rc = LNR_OK;
    auto done = [&](int code) {
        if (stage) cudaFreeHost(stage);
        fclose(f);
        if (code) lnr_index_destroy(ix); else *out = ix;
        return code;
    };
    if (cudaMallocHost(&stage, kIdxFileChunk) != cudaSuccess) return done(fail(ctx, LNR_E_CUDA, "cudaMallocHost failed in lnr_index_load"));
    uint64_t sa = 0, sb = 0;
    if (dtype)
    {
        ix->n_hs = hd.n_b;
        if (cudaMalloc(&ix->d_dir, (size_t)kDirSize * sizeof(i32)) != cudaSuccess || cudaMalloc(&ix->d_hs, (size_t)(hd.n_b + 8) * sizeof(u64)) != cudaSuccess ||
            cudaMalloc(&ix->d_hsy, (size_t)hd.n_b + 64) != cudaSuccess)
            return done(fail(ctx, LNR_E_CUDA, "cudaMalloc failed in lnr_index_load"));
        cudaMemsetAsync(ix->d_hs + hd.n_b, 0, 8 * sizeof(u64), ctx->stream);
        if ((rc = idx_file_stream(ctx, f, ix->d_dir, (size_t)kDirSize * sizeof(i32), false, stage, &sa))) return done(rc);
        if ((rc = idx_file_stream(ctx, f, ix->d_hs, (size_t)hd.n_b * sizeof(u64), false, stage, &sb))) return done(rc);
        if (sa != hd.sum_a || sb != hd.sum_b) return done(fail(ctx, LNR_E_ARG, "index file: checksum mismatch"));
        if (hd.n_b) k_idx_split_y<<<(u32)((hd.n_b + 255) / 256), 256, 0, ctx->stream>>>(ix->d_hs, hd.n_b, ix->d_hsy);
        if (index_build_dirx(ctx, ix) != cudaSuccess) return done(fail(ctx, LNR_E_CUDA, "cudaMalloc failed in lnr_index_load"));
    }
    else
    {
        ix->n_ysa = hd.n_a; ix->tab_len = hd.n_b; ix->empty_dir = hd.empty_dir; ix->n_dir_entries = hd.n_dir_entries;
        if (hd.empty_dir >= hd.n_a) return done(fail(ctx, LNR_E_ARG, "index file: inconsistent header"));
        if (cudaMalloc(&ix->d_ysa, (size_t)(hd.n_a + 8) * sizeof(u64)) != cudaSuccess || cudaMalloc(&ix->d_tab, (size_t)hd.n_b * sizeof(HNode)) != cudaSuccess)
            return done(fail(ctx, LNR_E_CUDA, "cudaMalloc failed in lnr_index_load"));
        cudaMemsetAsync(ix->d_ysa + hd.n_a, 0, 8 * sizeof(u64), ctx->stream);      // the seeding scan may look past the two terminators
        if ((rc = idx_file_stream(ctx, f, ix->d_ysa, (size_t)hd.n_a * sizeof(u64), false, stage, &sa))) return done(rc);
        if ((rc = idx_file_stream(ctx, f, ix->d_tab, (size_t)hd.n_b * sizeof(HNode), false, stage, &sb))) return done(rc);
        if (sa != hd.sum_a || sb != hd.sum_b) return done(fail(ctx, LNR_E_ARG, "index file: checksum mismatch"));
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return done(fail(ctx, LNR_E_CUDA, cudaGetErrorString(e)));
    return done(LNR_OK);
}
static int index_build_range(lnr_ctx * ctx, const lnr_genome * g, int index_type, unsigned threads_sem, u32 x_lo, u32 x_hi, lnr_index ** out)
{
    if (!ctx || !g || !out || threads_sem == 0) return LNR_E_ARG;
    if (index_type == 2)
    {
        if (x_lo != 0 || x_hi != (1u << kDirBits)) return fail(ctx, LNR_E_UNSUPPORTED, "lnr_index_build_shard (caller-side exchange) is DIndex only; use lnr_index_build_sharded for -i 2");
        return hindex_build(ctx, g, threads_sem, nullptr, out);
    }
    if (index_type != 1) return fail(ctx, LNR_E_UNSUPPORTED, "index_type must be 1 (DIndex, -i 1) or 2 (HIndex, -i 2)");
    return dindex_build(ctx, g, threads_sem, x_lo, x_hi, nullptr, out);
}
// ---- HIndex build (host orchestration) -------------------------------------------------------------------------------
// comm != nullptr: the X axis (18 bits) is cut into n_ranks ranges of equal pair count -- every rank derives the same cuts from
// the same per-X histogram of the whole genome --, a rank sorts and assembles only the pairs of its range, straight into its slice
// of the final ysa (blocks never straddle a cut: a block is one X), one all-gather of (pairs, blocks) tells everybody the
// displacements, one group of point-to-point transfers completes the ysa on every rank, and the directory is derived locally
// from the assembled ysa. The result does not depend on the number of ranks.
static int hindex_build(lnr_ctx * ctx, const lnr_genome * g, unsigned T, lnr_comm * comm, lnr_index ** out)
{
    cudaSetDevice(ctx->device);
    std::vector<HChunk> chunks;
    u64 n_samples = 0;
    for (uint32_t ci = 0; ci < g->n_contigs; ci++)
    {
        u64 len = g->len[ci];
        if (len < (u64)kSpanH + T) continue;
        u64 n = len - kSpanH + 1, q = n / T, r = n - q * T;
        for (unsigned t = 0; t < T; t++)   // index_util.cpp:742-760
        {
            HChunk ch;
            memset(&ch, 0, sizeof ch);
            ch.base_off = g->off[ci]; ch.len = (i64)len; ch.contig = ci;
            if (t < r) { ch.csize = (i64)q + 1; ch.start = (i64)((q + 1) * t); }
            else { ch.csize = (i64)q; ch.start = (i64)(len + 1 - kSpanH - q * (T - t)); }
            ch.k_first = (ch.start + kStepH - 1) / kStepH * kStepH;
            i64 k_end = ch.start + ch.csize;
            ch.n_samples = ch.k_first < k_end ? (k_end - ch.k_first + kStepH - 1) / kStepH : 0;
            if (ch.csize <= 0) continue;
            ch.sample0 = n_samples;
            n_samples += (u64)ch.n_samples;
            chunks.push_back(ch);
        }
    }
    if (!n_samples) return fail(ctx, LNR_E_UNSUPPORTED, "genome too small for the HIndex");
    lnr_index * ix = new lnr_index();
    ix->ctx = ctx; ix->index_type = 2; ix->d_dir = nullptr; ix->d_hs = nullptr; ix->n_hs = 0;
    HChunk * d_chunks = nullptr; u64 * d_body[2] = {nullptr, nullptr}; u32 * d_x[2] = {nullptr, nullptr};
    u32 * d_hist = nullptr; u64 * d_offs = nullptr; u64 * d_small = nullptr; u32 * d_flag = nullptr; u64 * d_bid = nullptr; u64 * d_starts = nullptr;
    u64 * d_goff = nullptr; u64 * d_glen = nullptr; u32 * d_xhist = nullptr;
    auto cleanup = [&]() {
        for (void * p : {(void *)d_chunks, (void *)d_body[0], (void *)d_body[1], (void *)d_x[0], (void *)d_x[1], (void *)d_hist, (void *)d_offs,
                         (void *)d_small, (void *)d_flag, (void *)d_bid, (void *)d_starts, (void *)d_goff, (void *)d_glen, (void *)d_xhist})
            if (p) cudaFree(p);
    };
#define CKH(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); cleanup(); lnr_index_destroy(ix); return LNR_E_CUDA; } } while (0)
    u32 n_chunks = (u32)chunks.size();
    CKH(cudaMalloc(&d_small, 64 * sizeof(u64)));
    CKH(cudaMemsetAsync(d_small, 0, 64 * sizeof(u64), ctx->stream));
    // ACGT only
    CKH(cudaMalloc(&d_goff, g->n_contigs * sizeof(u64)));
    CKH(cudaMalloc(&d_glen, g->n_contigs * sizeof(u64)));
    CKH(cudaMemcpyAsync(d_goff, g->off.data(), g->n_contigs * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    CKH(cudaMemcpyAsync(d_glen, g->len.data(), g->n_contigs * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    {
        LaunchScope ls(ctx, "k_check_acgt");
        k_check_acgt<<<ctx->n_sm * 8, 256, 0, ctx->stream>>>(g->d_bases, d_goff, d_glen, g->n_contigs, (u32 *)(d_small + 8));
    }
    CKH(cudaMalloc(&d_chunks, chunks.size() * sizeof(HChunk)));
    CKH(cudaMemcpyAsync(d_chunks, chunks.data(), chunks.size() * sizeof(HChunk), cudaMemcpyHostToDevice, ctx->stream));
    {
        LaunchScope ls(ctx, "k_hidx_prep");
        k_hidx_prep<<<(n_chunks + 63) / 64, 64, 0, ctx->stream>>>(g->d_bases, d_chunks, n_chunks);
    }
    u32 x_lo = 0, x_hi = kXRangeH;
    u64 pair_cap = n_samples;
    if (comm)
    {
        // pairs per X over the whole genome -> this rank's range [x_lo, x_hi) and its pair count (the arrays are sized by it)
        CKH(cudaMalloc(&d_xhist, (size_t)kXRangeH * sizeof(u32)));
        CKH(cudaMemsetAsync(d_xhist, 0, (size_t)kXRangeH * sizeof(u32), ctx->stream));
        {
            LaunchScope ls(ctx, "k_hidx_xhist");
            k_hidx_pairs<<<(u32)((n_samples + 255) / 256), 256, 0, ctx->stream>>>(g->d_bases, d_chunks, n_chunks, n_samples, nullptr, nullptr, nullptr,
                                                                                  0u, kXRangeH, d_xhist);
        }
        CKH(cudaGetLastError());
        std::vector<u32> xh(kXRangeH);
        CKH(cudaMemcpyAsync(xh.data(), d_xhist, (size_t)kXRangeH * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
        CKH(cudaStreamSynchronize(ctx->stream));
        std::vector<u32> cut((size_t)comm->n_ranks + 1, kXRangeH);
        lnr_hindex_shard_cuts(xh.data(), kXRangeH, comm->n_ranks, cut.data());
        x_lo = cut[(size_t)comm->rank]; x_hi = cut[(size_t)comm->rank + 1];
        pair_cap = 0;
        for (u32 x = x_lo; x < x_hi; x++) pair_cap += xh[x];
    }
    for (int i = 0; i < 2; i++) { CKH(cudaMalloc(&d_body[i], (size_t)(pair_cap + 8) * sizeof(u64))); CKH(cudaMalloc(&d_x[i], (size_t)(pair_cap + 8) * sizeof(u32))); }
    {
        LaunchScope ls(ctx, "k_hidx_pairs");
        k_hidx_pairs<<<(u32)((n_samples + 255) / 256), 256, 0, ctx->stream>>>(g->d_bases, d_chunks, n_chunks, n_samples, d_body[0], d_x[0],
                                                                              (unsigned long long *)d_small, x_lo, x_hi, nullptr);
    }
    CKH(cudaGetLastError());
    u64 h_small[16];
    CKH(cudaMemcpyAsync(h_small, d_small, sizeof h_small, cudaMemcpyDeviceToHost, ctx->stream));
    CKH(cudaStreamSynchronize(ctx->stream));
    const u64 n = h_small[0];                  // pairs of this build (of this rank's range)
    if (((u32 *)(h_small + 8))[0]) { cleanup(); lnr_index_destroy(ix); return fail(ctx, LNR_E_UNSUPPORTED, "HIndex (-i 2): genomes containing N are not supported"); }
    if (n > pair_cap) { cleanup(); lnr_index_destroy(ix); return fail(ctx, LNR_E_CUDA, "HIndex build: pair count exceeds the histogram's"); }
    // ---- sort: body descending (LSD over its varying bytes), then X ascending
    u64 init[4] = {0, ~0ULL, 0, ~0ULL};
    CKH(cudaMemcpyAsync(d_small + 16, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    {
        LaunchScope ls(ctx, "k_rs_masks");
        k_rs_masks<<<ctx->n_sm * 8, 256, 0, ctx->stream>>>(d_body[0], d_x[0], n, d_small + 16);
    }
    u64 masks[4];
    CKH(cudaMemcpyAsync(masks, d_small + 16, sizeof masks, cudaMemcpyDeviceToHost, ctx->stream));
    CKH(cudaStreamSynchronize(ctx->stream));
    const u64 vary_body = masks[0] ^ masks[1];
    const u32 vary_x = (u32)(masks[2] ^ masks[3]);
    const u32 n_tiles = (u32)((n + RS_TILE - 1) / RS_TILE);
    CKH(cudaMalloc(&d_hist, (size_t)256 * n_tiles * sizeof(u32) + 64));
    CKH(cudaMalloc(&d_offs, ((size_t)256 * n_tiles + STILE) * sizeof(u64)));
    int cur = 0;
    for (int pass = 0; pass < 11 && n; pass++)     // n == 0: a rank whose range is empty (tiny genome, many ranks)
    {
        RsDigit dg;
        if (pass < 8) { dg.from_aux = 0; dg.shift = 8 * pass; dg.xor_mask = ~0ULL; if (((vary_body >> (8 * pass)) & 255) == 0) continue; }
        else { dg.from_aux = 1; dg.shift = 8 * (pass - 8); dg.xor_mask = 0; if (((vary_x >> (8 * (pass - 8))) & 255) == 0) continue; }
        {
            LaunchScope ls(ctx, "k_rs_hist");
            k_rs_hist<<<n_tiles, RS_T, 0, ctx->stream>>>(d_body[cur], d_x[cur], n, dg, d_hist, n_tiles);
        }
        int rc = device_scan<u64>(ctx, d_hist, (u64)256 * n_tiles, 0, d_offs, d_small + 24, "k_scan_radix");
        if (rc) { cleanup(); lnr_index_destroy(ix); return rc; }
        {
            LaunchScope ls(ctx, "k_rs_scatter");
            k_rs_scatter<<<n_tiles, RS_T, 0, ctx->stream>>>(d_body[cur], d_x[cur], n, dg, d_offs, n_tiles, d_body[cur ^ 1], d_x[cur ^ 1]);
        }
        cur ^= 1;
    }
    CKH(cudaGetLastError());
    // ---- blocks
    CKH(cudaMalloc(&d_flag, (size_t)(n + STILE + 8) * sizeof(u32)));
    CKH(cudaMalloc(&d_bid, (size_t)(n + STILE + 8) * sizeof(u64)));
    {
        LaunchScope ls(ctx, "k_hidx_flags");
        k_hidx_flags<<<(u32)((n + 1 + 255) / 256), 256, 0, ctx->stream>>>(d_x[cur], n, d_flag);
    }
    {
        int rc = device_scan<u64>(ctx, d_flag, n + 1, 0, d_bid, d_small + 24, "k_scan_blocks");
        if (rc) { cleanup(); lnr_index_destroy(ix); return rc; }
    }
    u64 n_blocks = 0;
    CKH(cudaMemcpyAsync(&n_blocks, d_small + 24, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CKH(cudaStreamSynchronize(ctx->stream));
    // ---- sharded: (pairs, blocks) of every rank -> totals and the displacement of this rank's slice in the final ysa
    u64 n_all = n, nb_all = n_blocks, disp = 0;
    std::vector<u64> rank_words;                   // ysa words (pairs + blocks) of every rank's slice
    if (comm)
    {
        NcclApi & api = nccl_api();
        if (comm->n_ranks > 64) { cleanup(); lnr_index_destroy(ix); return fail(ctx, LNR_E_ARG, "more than 64 ranks"); }
        const u64 mine[2] = {n, n_blocks};
        std::vector<u64> cnts(2 * (size_t)comm->n_ranks);
        u64 * d_mine = d_small + 40; u64 * d_cnts = (u64 *)d_hist;        // the radix histogram is free by now (>= 64 bytes; see below)
        if (2 * (size_t)comm->n_ranks * sizeof(u64) > (size_t)256 * n_tiles * sizeof(u32) + 64)
        {
            cudaFree(d_hist); d_hist = nullptr;
            CKH(cudaMalloc(&d_hist, 2 * (size_t)comm->n_ranks * sizeof(u64)));
            d_cnts = (u64 *)d_hist;
        }
        CKH(cudaMemcpyAsync(d_mine, mine, sizeof mine, cudaMemcpyHostToDevice, ctx->stream));
        int nrc = api.AllGather(d_mine, d_cnts, 2, kNcclUint64, comm->comm, ctx->stream);
        cudaError_t ce = cudaMemcpyAsync(cnts.data(), d_cnts, cnts.size() * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
        if (nrc != 0 || ce != cudaSuccess) { cleanup(); lnr_index_destroy(ix); return fail(ctx, LNR_E_CUDA, "all-gather of the shard sizes failed"); }
        n_all = 0; nb_all = 0;
        for (int r = 0; r < comm->n_ranks; r++)
        {
            if (r == comm->rank) disp = n_all + nb_all;
            n_all += cnts[2 * (size_t)r]; nb_all += cnts[2 * (size_t)r + 1];
            rank_words.push_back(cnts[2 * (size_t)r] + cnts[2 * (size_t)r + 1]);
        }
    }
    if (n_all - nb_all <= 2) { cleanup(); lnr_index_destroy(ix); return fail(ctx, LNR_E_UNSUPPORTED, "genome too small for the HIndex (countMove <= 2, index_util.cpp:1333)"); }
    if (n_all + nb_all + 2 >= (1ULL << 32)) { cleanup(); lnr_index_destroy(ix); return fail(ctx, LNR_E_LIMIT, "ysa exceeds the 32-bit directory values (index_util.h: XNode val2)"); }
    CKH(cudaMalloc(&d_starts, (size_t)(n_blocks + 2) * sizeof(u64)));
    ix->n_ysa = n_all + nb_all + 2;
    ix->empty_dir = n_all + nb_all;
    CKH(cudaMalloc(&ix->d_ysa, (size_t)(ix->n_ysa + 8) * sizeof(u64)));
    CKH(cudaMemsetAsync(ix->d_ysa + n_all + nb_all, 0, 10 * sizeof(u64), ctx->stream));
    {
        LaunchScope ls(ctx, "k_hidx_starts");
        k_hidx_starts<<<(u32)((n + 1 + 255) / 256), 256, 0, ctx->stream>>>(d_flag, d_bid, n, d_starts);
    }
    if (n)
    {
        LaunchScope ls(ctx, "k_hidx_write");
        k_hidx_write<<<(u32)((n + 255) / 256), 256, 0, ctx->stream>>>(d_body[cur], d_x[cur], d_flag, d_bid, d_starts, n, ix->d_ysa + disp);
    }
    CKH(cudaGetLastError());
    if (comm && comm->n_ranks > 1)
    {
        // the one exchange step: every rank's slice of ysa, in place at its displacement
        NcclApi & api = nccl_api();
        int nrc = 0;
        {
            LaunchScope ls(ctx, "nccl_exchange", 0);
            nrc |= api.GroupStart();
            u64 d = 0;
            for (int r = 0; r < comm->n_ranks; r++)
            {
                if (r != comm->rank)
                {
                    if (rank_words[(size_t)r]) nrc |= api.Recv(ix->d_ysa + d, (size_t)rank_words[(size_t)r], kNcclUint64, r, comm->comm, ctx->stream);
                    if (n + n_blocks) nrc |= api.Send(ix->d_ysa + disp, (size_t)(n + n_blocks), kNcclUint64, r, comm->comm, ctx->stream);
                }
                d += rank_words[(size_t)r];
            }
            nrc |= api.GroupEnd();
        }
        if (nrc != 0) { cleanup(); lnr_index_destroy(ix); return fail(ctx, LNR_E_CUDA, "NCCL exchange of the ysa slices failed"); }
    }
    // ---- directory: from this build's block starts, or (sharded) from the head words of the assembled ysa
    const u64 n_words = n_all + nb_all;
    CKH(cudaMemsetAsync(d_small + 32, 0, sizeof(u64), ctx->stream));
    {
        LaunchScope ls(ctx, "k_hidx_dir_count");
        if (comm) k_hidx_dir_scan<<<(u32)((n_words + 255) / 256), 256, 0, ctx->stream>>>(ix->d_ysa, n_words, 0, (unsigned long long *)(d_small + 32), nullptr, 0);
        else k_hidx_dir<<<(u32)((n_blocks + 255) / 256), 256, 0, ctx->stream>>>(ix->d_ysa, d_starts, n_blocks, 0, (unsigned long long *)(d_small + 32), nullptr, 0);
    }
    u64 n_ent = 0;
    CKH(cudaMemcpyAsync(&n_ent, d_small + 32, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CKH(cudaStreamSynchronize(ctx->stream));
    u64 tl = 1;
    while ((float)tl < (float)n_ent * 1.6f) tl <<= 1;   // XString::_fullSize index_util.cpp:221, alpha 1.6
    ix->tab_len = tl; ix->n_dir_entries = n_ent;
    CKH(cudaMalloc(&ix->d_tab, (size_t)tl * sizeof(HNode)));
    CKH(cudaMemsetAsync(ix->d_tab, 0, (size_t)tl * sizeof(HNode), ctx->stream));
    {
        LaunchScope ls(ctx, "k_hidx_dir_insert");
        if (comm) k_hidx_dir_scan<<<(u32)((n_words + 255) / 256), 256, 0, ctx->stream>>>(ix->d_ysa, n_words, 1, nullptr, ix->d_tab, tl - 1);
        else k_hidx_dir<<<(u32)((n_blocks + 255) / 256), 256, 0, ctx->stream>>>(ix->d_ysa, d_starts, n_blocks, 1, nullptr, ix->d_tab, tl - 1);
    }
    CKH(cudaGetLastError());
    CKH(cudaStreamSynchronize(ctx->stream));
    cleanup();
#undef CKH
    *out = ix;
    return LNR_OK;
}

// comm == nullptr: the buckets [x_lo, x_hi) only (whole index, or one shard for a caller-side exchange).
// comm != nullptr: this rank builds [x_lo, x_hi) straight into its slice of the final arrays, then one grouped exchange
// (every rank broadcasts its hs slice and its dir slice in place, at their displacements) completes them on every rank.
static int dindex_build(lnr_ctx * ctx, const lnr_genome * g, unsigned threads_sem, u32 x_lo, u32 x_hi, lnr_comm * comm, lnr_index ** out)
{
    // with a communicator the rank's bucket range is decided below, from the minimizer histogram every rank computes
    // identically: equal-width ranges are badly unbalanced (minimizers crowd towards small X: the lower half of the range
    // holds 99.6 % of the records of a random genome)
    std::vector<u32> rank_lo;   // comm: first bucket of every rank's range, n_ranks + 1 entries
    if (comm)
        for (int r = 0; r <= comm->n_ranks; r++) rank_lo.push_back((u32)(((u64)r << kDirBits) / (u64)comm->n_ranks));   // empty genome: any split does
    cudaSetDevice(ctx->device);
    // chunk table (createDIndex :1654-1670)
    std::vector<IdxChunk> chunks;
    std::vector<u32> tile0;
    u64 n_samples = 0;
    u32 n_tiles = 0;
    for (uint32_t ci = 0; ci < g->n_contigs; ci++)
        for (unsigned c = 0; c < threads_sem; c++)
        {
            IdxChunk ch;
            memset(&ch, 0, sizeof ch);
            ch.base_off = g->off[ci]; ch.len = (i64)g->len[ci]; ch.contig = ci;
            idx_chunk_range(ch.len, threads_sem, c, ch.t_str, ch.n_samples);
            if (ch.n_samples <= 0) continue;
            ch.sample0 = n_samples;
            n_samples += (u64)ch.n_samples;
            tile0.push_back(n_tiles);
            n_tiles += (u32)((ch.n_samples + ITILE - 1) / ITILE);
            chunks.push_back(ch);
        }
    lnr_index * ix = new lnr_index();
    ix->ctx = ctx; ix->index_type = 1; ix->d_dir = nullptr; ix->d_hs = nullptr; ix->n_hs = 0;
    auto cleanup = [&]() {};
#define CKI(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); cleanup(); lnr_index_destroy(ix); return LNR_E_CUDA; } } while (0)
    const bool full_range = x_lo == 0 && x_hi >= (u32)(kDirSize - 1);
    const u32 n_chunks = (u32)chunks.size();
    const u64 n_slots = (u64)n_tiles * ITILE;
    // every temporary of the build is carved out of the context's scratch arena (the per-warp arenas of the mapping
    // kernels use the same allocation later): a second build, or a build after a batch, pays no cudaMalloc / cudaFree
    size_t off_t = 0;
    auto carve = [&](size_t bytes) { size_t o = off_t; off_t += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_cnt = carve(((size_t)kDirSize + 16) * sizeof(u32));
    const size_t o_small = carve(4 * IPARTS * sizeof(u64) + ICOARSE * sizeof(u32) + 4 * (IPARTS + 1) * sizeof(u32));
    const size_t o_chunks = carve(std::max<size_t>(chunks.size(), 1) * sizeof(IdxChunk));
    const size_t o_tile0 = carve(std::max<size_t>(tile0.size(), 1) * sizeof(u32));
    const size_t o_px0 = carve((size_t)(n_slots + 16) * sizeof(u32)), o_pr0 = carve((size_t)(n_slots + 16) * sizeof(u64));
    const size_t o_px1 = carve((size_t)(n_slots + 16) * sizeof(u32)), o_pr1 = carve((size_t)(n_slots + 16) * sizeof(u64));
    CKI(ctx->arena.reserve(off_t));
    u8 * T0 = ctx->arena.as<u8>();
    u32 * d_cnt = (u32 *)(T0 + o_cnt);
    u64 * d_small = (u64 *)(T0 + o_small);
    IdxChunk * d_chunks = (IdxChunk *)(T0 + o_chunks);
    u32 * d_tile0 = (u32 *)(T0 + o_tile0);
    u32 * d_px[2] = {(u32 *)(T0 + o_px0), (u32 *)(T0 + o_px1)};
    u64 * d_pr[2] = {(u64 *)(T0 + o_pr0), (u64 *)(T0 + o_pr1)};
    CKI(cudaMalloc(&ix->d_dir, (size_t)kDirSize * sizeof(i32)));
    // d_small: [0] scan total, [IPARTS..) partition counts, [2 IPARTS..) offsets, [3 IPARTS..) fill, the coarse histogram, then
    // the split points and dir at the split points
    CKI(cudaMemsetAsync(d_cnt, 0, ((size_t)kDirSize + 16) * sizeof(u32), ctx->stream));
    CKI(cudaMemsetAsync(d_small, 0, 4 * IPARTS * sizeof(u64) + ICOARSE * sizeof(u32), ctx->stream));
    u64 * d_total = d_small;
    unsigned long long * d_part_cnt = (unsigned long long *)(d_small + IPARTS);
    u64 * d_part_off = d_small + 2 * IPARTS;
    unsigned long long * d_part_fill = (unsigned long long *)(d_small + 3 * IPARTS);
    u32 * d_coarse = (u32 *)(d_small + 4 * IPARTS);
    u32 * d_splitx = d_coarse + ICOARSE;                 // 2 (IPARTS + 1) bucket indices: xa, xb of every partition
    i32 * d_splitdir = (i32 *)(d_splitx + 2 * (IPARTS + 1));
    IdxParts parts;
    u64 part_off[IPARTS + 1];
    for (int q = 0; q <= IPARTS; q++) { parts.split[q] = q == IPARTS ? (1u << kDirBits) : 0; part_off[q] = 0; }
    u64 n_pairs = 0;
    if (n_chunks)
    {
        CKI(cudaMemcpyAsync(d_chunks, chunks.data(), chunks.size() * sizeof(IdxChunk), cudaMemcpyHostToDevice, ctx->stream));
        CKI(cudaMemcpyAsync(d_tile0, tile0.data(), tile0.size() * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
        {
            LaunchScope ls(ctx, "k_idx_prep");
            k_idx_prep<<<(n_chunks + 127) / 128, 128, 0, ctx->stream>>>(g->d_bases, d_chunks, n_chunks);
        }
        {
            LaunchScope ls(ctx, "k_idx_emit");
            k_idx_emit<<<n_tiles, IT, 0, ctx->stream>>>(g->d_bases, d_chunks, d_tile0, n_chunks, d_px[0], d_pr[0], comm ? 0u : x_lo,
                                                        comm ? (1u << kDirBits) : x_hi, d_coarse);
        }
        CKI(cudaGetLastError());
        // splitters of ~equal record count from the sampled coarse histogram (at least one coarse bin per partition)
        std::vector<u32> coarse(ICOARSE);
        CKI(cudaMemcpyAsync(coarse.data(), d_coarse, ICOARSE * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
        CKI(cudaStreamSynchronize(ctx->stream));
        // equal-count cut points of [bin_lo, bin_hi) at coarse-bin granularity: out[0] = first bucket, out[n] = last + 1
        auto equal_cuts = [&](int bin_lo, int bin_hi, int n, u32 * cut) {
            u64 csum = 0;
            for (int b = bin_lo; b < bin_hi; b++) csum += coarse[(size_t)b];
            cut[0] = (u32)bin_lo << ICOARSE_SHIFT;
            u64 run = 0; int q = 1;
            for (int bin = bin_lo; bin < bin_hi && q < n; bin++)
            {
                run += coarse[(size_t)bin];
                while (q < n && csum && run * (u64)n >= csum * (u64)q) cut[q++] = (u32)(bin + 1) << ICOARSE_SHIFT;
            }
            for (; q <= n; q++) cut[q] = (u32)bin_hi << ICOARSE_SHIFT;   // nothing sampled beyond: empty ranges
            cut[n] = (u32)bin_hi << ICOARSE_SHIFT;
            for (q = 1; q <= n; q++) if (cut[q] < cut[q - 1]) cut[q] = cut[q - 1];
        };
        if (comm)
        {
            rank_lo.assign((size_t)comm->n_ranks + 1, 0);
            equal_cuts(0, ICOARSE, comm->n_ranks, rank_lo.data());
            x_lo = rank_lo[(size_t)comm->rank]; x_hi = rank_lo[(size_t)comm->rank + 1];
        }
        // the 64 partitions of this build's own range
        {
            const int bin_lo = (int)(x_lo >> ICOARSE_SHIFT), bin_hi = (int)(((u64)x_hi + (1u << ICOARSE_SHIFT) - 1) >> ICOARSE_SHIFT);
            equal_cuts(bin_lo, std::max(bin_hi, bin_lo), IPARTS, parts.split);
            parts.split[0] = 0; parts.split[IPARTS] = 1u << kDirBits;   // the outer partitions absorb whatever lies outside (nothing does)
        }
        {
            LaunchScope ls(ctx, "k_idx_partcount");
            k_idx_partcount<<<ctx->n_sm * 8, 256, 0, ctx->stream>>>(d_px[0], n_slots, parts, d_part_cnt, x_lo, x_hi);
        }
        u64 h_cnt[IPARTS];
        CKI(cudaMemcpyAsync(h_cnt, d_part_cnt, sizeof h_cnt, cudaMemcpyDeviceToHost, ctx->stream));
        CKI(cudaStreamSynchronize(ctx->stream));
        for (int q = 0; q < IPARTS; q++) part_off[q + 1] = part_off[q] + h_cnt[q];
        n_pairs = part_off[IPARTS];      // emitted samples, including those of buckets that will turn out to be omitted
        if (n_pairs)
        {
            CKI(cudaMemcpyAsync(d_part_off, part_off, IPARTS * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
            {
                LaunchScope ls(ctx, "k_idx_part");
                k_idx_part<<<(u32)((n_slots + PTILE - 1) / PTILE), PT, 0, ctx->stream>>>(d_px[0], d_pr[0], n_slots, parts, d_part_off, d_part_fill, d_px[1], d_pr[1],
                                                                                         x_lo, x_hi);
            }
            {
                LaunchScope ls(ctx, "k_idx_count");
                k_idx_count<<<(u32)((n_pairs + 1023) / 1024), 256, 0, ctx->stream>>>(d_px[1], n_pairs, d_cnt);
            }
            CKI(cudaGetLastError());
        }
    }
    // omit buckets > 400, exclusive prefix sum over 2^26+1 entries (:1702-1721)
    {
        int rc = device_scan<i32>(ctx, d_cnt, kDirSize, kIdxOmit, ix->d_dir, d_total, "k_scan_dir");
        if (rc) { cleanup(); lnr_index_destroy(ix); return rc; }
    }
    CKI(cudaMemsetAsync(d_cnt, 0, ((size_t)kDirSize + 16) * sizeof(u32), ctx->stream));   // from here on: the fill counters of k_idx_place
    // where every partition's buckets start and end in hs (this build's own coordinates)
    u32 h_splitx[2 * (IPARTS + 1)];
    i32 h_splitdir[2 * (IPARTS + 1)];
    for (int q = 0; q < IPARTS; q++)
    {
        const u32 xa = std::max(parts.split[q], x_lo), xb = std::min(parts.split[q + 1], std::min(x_hi, (u32)(kDirSize - 1)));
        h_splitx[2 * q] = xa; h_splitx[2 * q + 1] = xb > xa ? xb : xa;
    }
    h_splitx[2 * IPARTS] = h_splitx[2 * IPARTS + 1] = 0;
    CKI(cudaMemcpyAsync(d_splitx, h_splitx, sizeof h_splitx, cudaMemcpyHostToDevice, ctx->stream));
    k_gather_i32<<<1, 2 * (IPARTS + 1), 0, ctx->stream>>>(ix->d_dir, d_splitx, 2 * (IPARTS + 1), d_splitdir);
    u64 h_small[1];
    CKI(cudaMemcpyAsync(h_small, d_small, sizeof h_small, cudaMemcpyDeviceToHost, ctx->stream));
    CKI(cudaMemcpyAsync(h_splitdir, d_splitdir, sizeof h_splitdir, cudaMemcpyDeviceToHost, ctx->stream));
    CKI(cudaStreamSynchronize(ctx->stream));
    const u64 local_total = h_small[0];
    u64 total = local_total, hs_base = 0;            // this build's records start at hs_base of the (final) hs
    std::vector<u64> rank_cnt;
    if (comm)
    {
        // record counts of all ranks -> displacements of the slices in the final hs
        NcclApi & api = nccl_api();
        rank_cnt.assign((size_t)comm->n_ranks, 0);
        u64 * d_cnts = (u64 *)d_part_fill;            // 64 free words by now
        if (comm->n_ranks > IPARTS) { cleanup(); lnr_index_destroy(ix); return fail(ctx, LNR_E_ARG, "more than 64 ranks"); }
        int nrc = api.AllGather(d_total, d_cnts, 1, kNcclUint64, comm->comm, ctx->stream);
        cudaError_t ce = cudaMemcpyAsync(rank_cnt.data(), d_cnts, (size_t)comm->n_ranks * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
        if (nrc != 0 || ce != cudaSuccess) { cleanup(); lnr_index_destroy(ix); return fail(ctx, LNR_E_CUDA, "all-gather of the shard sizes failed"); }
        total = 0;
        for (int r = 0; r < comm->n_ranks; r++) { if (r == comm->rank) hs_base = total; total += rank_cnt[(size_t)r]; }
    }
    if (total >= (1ULL << 31)) { cleanup(); lnr_index_destroy(ix); return fail(ctx, LNR_E_LIMIT, "hs exceeds int32 bucket offsets (index_util.h:101)"); }
    ix->n_hs = total;
    CKI(cudaMalloc(&ix->d_hs, (size_t)(total + 8) * sizeof(u64)));
    const bool derive_here = !comm;                  // sharded: Y bytes and lookup sectors are derived after the exchange
    CKI(cudaMalloc(&ix->d_hsy, (size_t)total + 64));
    if (full_range || comm) CKI(cudaMalloc(&ix->d_dirx, (size_t)(kDirSize - 1) * 16 * kDirxQuads));
    {
        // partition by partition: stage the records in their buckets, then move every record to its rank -- while the
        // partition's range is still in L2. The pair buffers of the emit pass are free by now and serve as the staging area.
        LaunchScope ls(ctx, "k_idx_place_rank", 2 * IPARTS);
        u64 * d_tmp = d_pr[0]; u32 * d_tmpx = d_px[0];
        for (int q = 0; q < IPARTS; q++)
        {
            const u64 np = part_off[q + 1] - part_off[q];
            if (!np) continue;
            k_idx_place<<<(u32)((np + 1023) / 1024), 256, 0, ctx->stream>>>(d_px[1] + part_off[q], d_pr[1] + part_off[q], np, ix->d_dir, d_cnt, d_tmp, d_tmpx);
            const u64 p0 = (u64)h_splitdir[2 * q], p1 = (u64)h_splitdir[2 * q + 1];
            if (p1 > p0)
                k_idx_rank<<<(u32)((p1 - p0 + 255) / 256), 256, 0, ctx->stream>>>(d_tmp, d_tmpx, p0, p1 - p0, ix->d_dir, ix->d_hs + hs_base,
                                                                                  derive_here ? ix->d_hsy : nullptr);
        }
    }
    CKI(cudaGetLastError());
    if (derive_here && full_range)
    {
        LaunchScope ls(ctx, "k_idx_dirx");
        k_idx_dirx<<<((kDirSize - 1) + 255) / 256, 256, 0, ctx->stream>>>(ix->d_dir, ix->d_hsy, kDirSize - 1, ix->d_dirx);
    }
    if (comm)
    {
        if (hs_base)
        {
            // dir of this rank's buckets, and the one entry behind them, in final coordinates
            LaunchScope ls(ctx, "k_idx_rebase");
            const u32 n = x_hi - x_lo + 1;
            k_idx_rebase<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ix->d_dir, x_lo, n, (i32)hs_base);
        }
        // the one exchange step: every rank's hs slice and dir slice, in place at their displacements (no padding, no staging
        // copy). One group of point-to-point transfers -- every rank sends its two slices to every peer and receives theirs --
        // so all pairs move at once over NVSwitch (a group of broadcasts with different roots runs them one after the other)
        NcclApi & api = nccl_api();
        int nrc = 0;
        {
            LaunchScope ls(ctx, "nccl_exchange", 0);
            nrc |= api.GroupStart();
            const u32 my_lo = rank_lo[(size_t)comm->rank], my_hi = rank_lo[(size_t)comm->rank + 1];
            const size_t my_dir = (size_t)(my_hi - my_lo) + (comm->rank == comm->n_ranks - 1 ? 1 : 0);
            u64 disp = 0;
            for (int r = 0; r < comm->n_ranks; r++)
            {
                const u32 lo = rank_lo[(size_t)r], hi = rank_lo[(size_t)r + 1];
                const size_t n_dir = (size_t)(hi - lo) + (r == comm->n_ranks - 1 ? 1 : 0);
                if (r != comm->rank)
                {
                    if (rank_cnt[(size_t)r]) nrc |= api.Recv(ix->d_hs + disp, (size_t)rank_cnt[(size_t)r], kNcclUint64, r, comm->comm, ctx->stream);
                    if (n_dir) nrc |= api.Recv(ix->d_dir + lo, n_dir, kNcclInt32, r, comm->comm, ctx->stream);
                    if (local_total) nrc |= api.Send(ix->d_hs + hs_base, (size_t)local_total, kNcclUint64, r, comm->comm, ctx->stream);
                    if (my_dir) nrc |= api.Send(ix->d_dir + my_lo, my_dir, kNcclInt32, r, comm->comm, ctx->stream);
                }
                disp += rank_cnt[(size_t)r];
            }
            nrc |= api.GroupEnd();
        }
        if (nrc != 0) { cleanup(); lnr_index_destroy(ix); return fail(ctx, LNR_E_CUDA, "NCCL exchange of the index slices failed"); }
        if (total)
        {
            LaunchScope ls(ctx, "k_idx_split_y");
            k_idx_split_y<<<(u32)((total + 255) / 256), 256, 0, ctx->stream>>>(ix->d_hs, total, ix->d_hsy);
        }
        {
            LaunchScope ls(ctx, "k_idx_dirx");
            k_idx_dirx<<<((kDirSize - 1) + 255) / 256, 256, 0, ctx->stream>>>(ix->d_dir, ix->d_hsy, kDirSize - 1, ix->d_dirx);
        }
        CKI(cudaGetLastError());
    }
    CKI(cudaStreamSynchronize(ctx->stream));
    cleanup();
#undef CKI
    *out = ix;
    return LNR_OK;
}
int lnr_index_export_dindex(const lnr_index * ix, int32_t * dir, uint64_t * hs, uint64_t hs_cap, uint64_t * n_hs)
{
    if (!ix || ix->index_type != 1) return LNR_E_ARG;
    lnr_ctx * ctx = ix->ctx;
    cudaSetDevice(ctx->device);
    if (n_hs) *n_hs = ix->n_hs;
    if (dir) CK(cudaMemcpy(dir, ix->d_dir, (size_t)kDirSize * sizeof(i32), cudaMemcpyDeviceToHost));
    if (hs)
    {
        if (hs_cap < ix->n_hs) return fail(ctx, LNR_E_CAPACITY, "hs buffer too small");
        CK(cudaMemcpy(hs, ix->d_hs, (size_t)ix->n_hs * sizeof(u64), cudaMemcpyDeviceToHost));
    }
    return LNR_OK;
}
void lnr_index_destroy(lnr_index * ix)
{
    if (!ix) return;
    cudaSetDevice(ix->ctx->device);
    if (ix->d_dir) cudaFree(ix->d_dir);
    if (ix->d_hs) cudaFree(ix->d_hs);
    if (ix->d_hsy) cudaFree(ix->d_hsy);
    if (ix->d_dirx) cudaFree(ix->d_dirx);
    if (ix->d_ysa) cudaFree(ix->d_ysa);
    if (ix->d_tab) cudaFree(ix->d_tab);
    delete ix;
}
int lnr_index_export_hindex(const lnr_index * ix, uint64_t * ysa, uint64_t ysa_cap, uint64_t * n_ysa, uint64_t * keyvals, uint64_t kv_cap,
                            uint64_t * n_kv, uint64_t * empty_dir, uint64_t * table_len)
{
    if (!ix || ix->index_type != 2) return LNR_E_ARG;
    lnr_ctx * ctx = ix->ctx;
    cudaSetDevice(ctx->device);
    if (n_ysa) *n_ysa = ix->n_ysa;
    if (n_kv) *n_kv = ix->n_dir_entries;
    if (empty_dir) *empty_dir = ix->empty_dir;
    if (table_len) *table_len = ix->tab_len;
    if (ysa)
    {
        if (ysa_cap < ix->n_ysa) return fail(ctx, LNR_E_CAPACITY, "ysa buffer too small");
        CK(cudaMemcpy(ysa, ix->d_ysa, (size_t)ix->n_ysa * sizeof(u64), cudaMemcpyDeviceToHost));
    }
    if (keyvals)
    {
        if (kv_cap < ix->n_dir_entries) return fail(ctx, LNR_E_CAPACITY, "keyvals buffer too small");
        u64 * d_kv = nullptr; unsigned long long * d_n = nullptr;
        CK(cudaMalloc(&d_kv, (size_t)(ix->n_dir_entries + 1) * 16));
        CK(cudaMalloc(&d_n, 8));
        CK(cudaMemset(d_n, 0, 8));
        k_hidx_dir_export<<<(u32)((ix->tab_len + 255) / 256), 256>>>(ix->d_tab, ix->tab_len, d_kv, d_n);
        std::vector<std::pair<u64, u64> > kv(ix->n_dir_entries);
        cudaError_t e = cudaMemcpy(kv.data(), d_kv, (size_t)ix->n_dir_entries * 16, cudaMemcpyDeviceToHost);
        cudaFree(d_kv); cudaFree(d_n);
        if (e != cudaSuccess) return fail(ctx, LNR_E_CUDA, cudaGetErrorString(e));
        std::sort(kv.begin(), kv.end());   // the physical layout is not part of the contract
        for (size_t i = 0; i < kv.size(); i++) { keyvals[2 * i] = kv[i].first; keyvals[2 * i + 1] = kv[i].second; }
    }
    return LNR_OK;
}

// ---- seeding pass (count / scan / fill) -----------------------------------------------------------------------
// tasks: host copy (sample0/n_samples filled); d_tasks: device copy. On return anchorsA holds the anchors,
// *d_aoff the per-sample offsets (n_samples + 1 entries, in sample_info's tail buffer), total anchors in *total.
// the most raw anchors any single task of the pass has: what the big-arena launches (one task per warp) must be able to hold
__global__ void k_task_max_anchors(const SeedTask * __restrict__ tasks, u32 n_tasks, const u64 * __restrict__ aoff, unsigned long long * out)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    u64 m = 0;
    if (i < n_tasks) { const SeedTask t = tasks[i]; m = aoff[t.sample0 + t.n_samples] - aoff[t.sample0]; }
    for (int o = 16; o; o >>= 1) { u64 v = __shfl_xor_sync(0xffffffffu, m, o); m = v > m ? v : m; }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, (unsigned long long)m);
}
static int seeding_pass(lnr_ctx * ctx, const lnr_index * ix, const u8 * d_bases, const u64 * d_read_off, SeedTask * d_tasks,
                        u32 n_tasks, u64 n_samples, DevBuf & aoff_buf, u64 * total_out, const char * tag, u64 * max_task_out = nullptr)
{
    if (ix->index_type == 1 && !ix->d_dirx)
    {
        // an index built as one shard of a sharded build and used on its own: finish its lookup table now
        static std::mutex mu;
        std::lock_guard<std::mutex> lk(mu);
        CK(index_build_dirx(ctx, const_cast<lnr_index *>(ix)));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    const bool hx_mode = ix->index_type == 2;
    if (hx_mode) CK(ctx->sample_info.reserve((size_t)(n_samples + 1) * sizeof(u64)));
    CK(ctx->sample_cnt.reserve((size_t)(n_samples + STILE + 1) * sizeof(u32)));
    CK(aoff_buf.reserve((size_t)(n_samples + STILE + 1) * sizeof(u64)));
    CK(ctx->misc.reserve(1024));
    // match entries of the count pass (k_seed_count): 4 per sample on average is twice what this workload produces; a warp
    // that finds its pool full is re-scanned by the fill pass
    u32 list_cap = (u32)std::min<u64>(4 * n_samples + (1u << 20), 0xfffffff0ull);
    if (const char * e = getenv("LNR_MASK_WORDS")) { long v = atol(e); if (v >= (long)kMaskPools && (u64)v < list_cap) list_cap = (u32)v; }   // tests: force the re-scan path
    const u64 n_swarps = (n_samples + 31) / 32;
    if (!hx_mode)
    {
        CK(ctx->seed_masks.reserve((size_t)list_cap * sizeof(u32)));
        CK(ctx->seed_mask_off.reserve((size_t)(n_swarps + 1) * sizeof(SeedWarpRec)));
        CK(ctx->task_info.reserve((size_t)(n_tasks + 1) * sizeof(SeedTaskInfo)));
    }
    CK(ctx->mask_ctr.reserve(kMaskPools * sizeof(unsigned int)));
    unsigned int * d_pool_ctr = ctx->mask_ctr.as<unsigned int>();
    CK(cudaMemsetAsync(d_pool_ctr, 0, kMaskPools * sizeof(unsigned int), ctx->stream));
    const u32 pool_cap = list_cap / kMaskPools;
    u64 * d_total = ctx->misc.as<u64>();
    unsigned long long * d_counters = (unsigned long long *)(ctx->misc.as<u64>() + 8);
    HIndexDev hx = {ix->d_ysa, ix->n_ysa, ix->empty_dir, ix->d_tab, ix->tab_len ? ix->tab_len - 1 : 0};
    {
        LaunchScope ls(ctx, "k_seed_prep");
        if (hx_mode) k_hseed_prep<<<(n_tasks + 127) / 128, 128, 0, ctx->stream>>>(d_bases, d_read_off, d_tasks, n_tasks);
        else k_seed_prep<<<(n_tasks + 127) / 128, 128, 0, ctx->stream>>>(d_bases, d_read_off, d_tasks, n_tasks, ctx->task_info.as<SeedTaskInfo>(),
                                                                         ctx->seed_mask_off.as<SeedWarpRec>());
    }
    // one extra zero count so that aoff[n_samples] = total
    CK(cudaMemsetAsync(ctx->sample_cnt.as<u32>() + n_samples, 0, sizeof(u32), ctx->stream));
    if (n_samples)
    {
        LaunchScope ls(ctx, tag);
        if (hx_mode)
            k_hseed_count<<<(u32)((n_samples + 255) / 256), 256, 0, ctx->stream>>>(d_bases, d_read_off, d_tasks, n_tasks, n_samples, hx,
                                                                                   ctx->sample_info.as<u64>(), ctx->sample_cnt.as<u32>(), d_counters);
        else
            k_seed_count<<<(u32)((n_samples + 255) / 256), 256, 0, ctx->stream>>>(d_bases, d_read_off, d_tasks, ctx->task_info.as<SeedTaskInfo>(), n_tasks, n_samples, ix->d_dirx,
                                                                                  ix->d_hsy, ctx->sample_cnt.as<u32>(), ctx->seed_mask_off.as<SeedWarpRec>(),
                                                                                  ctx->seed_masks.as<u32>(), pool_cap, d_pool_ctr);
    }
    CK(cudaGetLastError());
    int rc = device_scan<u64>(ctx, ctx->sample_cnt.as<u32>(), n_samples + 1, 0, aoff_buf.as<u64>(), d_total, "k_scan_seeds");
    if (rc) return rc;
    CK(cudaMemsetAsync(d_total + 1, 0, sizeof(u64), ctx->stream));
    if (n_tasks) k_task_max_anchors<<<(n_tasks + 255) / 256, 256, 0, ctx->stream>>>(d_tasks, n_tasks, aoff_buf.as<u64>(), (unsigned long long *)(d_total + 1));
    u64 tm[2] = {0, 0};
    CK(cudaMemcpyAsync(tm, d_total, 2 * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const u64 total = tm[0];
    *total_out = total;
    if (max_task_out) *max_task_out = tm[1];
    size_t need = (size_t)(total + n_tasks + 8) * sizeof(u64);
    CK(ctx->anchorsA.reserve(need));
    CK(ctx->anchorsB.reserve(need));
    if (n_samples)
    {
        LaunchScope ls(ctx, std::string(tag) == "k_seed_count" ? "k_seed_fill" : "k_seed_fill_remap");
        if (hx_mode)
            k_hseed_fill<<<(u32)((n_samples + 255) / 256), 256, 0, ctx->stream>>>(d_read_off, d_tasks, n_tasks, n_samples, hx,
                                                                                  ctx->sample_info.as<u64>(), aoff_buf.as<u64>(), ctx->anchorsA.as<u64>());
        else
            k_seed_fill<<<(u32)((n_samples + 255) / 256), 256, 0, ctx->stream>>>(d_bases, d_read_off, d_tasks, ctx->task_info.as<SeedTaskInfo>(), n_tasks, n_samples, ix->d_hs, ix->d_dirx,
                                                                                 aoff_buf.as<u64>(), ctx->anchorsA.as<u64>(), ctx->sample_cnt.as<u32>(),
                                                                                 ctx->seed_mask_off.as<SeedWarpRec>(), ctx->seed_masks.as<u32>(), d_counters);
    }
    if (n_samples && !hx_mode)
    {
        LaunchScope ls(ctx, "k_seed_stats");
        k_seed_stats<<<ctx->n_sm * 2, 256, 0, ctx->stream>>>(ctx->seed_mask_off.as<SeedWarpRec>(), n_swarps, d_counters);
        ctx->anchors_host_total += total;     // A: the scan's total (no device pass needed)
    }
    CK(cudaGetLastError());
    return LNR_OK;
}

// ---- the batch ---------------------------------------------------------------------------------------------------
// One bulk host->device upload at a time per process: concurrent callers (the reference calls p_calRecords from -t
// threads, each with its own lnr_ctx) would otherwise upload simultaneously at half the PCIe rate each and then all
// compute simultaneously -- lockstep, copy engine and SMs idle in turn. Serialising the uploads staggers the callers so
// that one thread's upload overlaps the others' kernels.
static std::mutex g_upload_mutex;

__global__ void k_stage_copy(const u8 * __restrict__ src, u8 * __restrict__ dst, size_t bytes)
{
    size_t n16 = bytes / 16;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    for (size_t k = i; k < n16; k += step) ((uint4 *)dst)[k] = ((const uint4 *)src)[k];
    for (size_t k = n16 * 16 + i; k < bytes; k += step) dst[k] = src[k];
}
// host table -> device through the pinned staging area (see HostStage); falls back to the copy engine when the staging
// area is full
static cudaError_t upload_small(lnr_ctx * ctx, void * dst, const void * src, size_t bytes)
{
    if (!bytes) return cudaSuccess;
    void * st = ctx->stage.take(bytes);
    if (!st) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
    memcpy(st, src, bytes);
    unsigned blocks = (unsigned)std::min<size_t>((bytes / 16 + 255) / 256 + 1, 296);
    k_stage_copy<<<blocks, 256, 0, ctx->stream>>>((const u8 *)st, (u8 *)dst, bytes);
    return cudaGetLastError();
}

struct HostTrace   // LNR_TRACE=1: host wall-clock of the phases of one batch (stderr)
{
    bool on; std::chrono::steady_clock::time_point t0; cudaStream_t st; std::string log;
    HostTrace(cudaStream_t s) : on(getenv("LNR_TRACE") != nullptr), t0(std::chrono::steady_clock::now()), st(s) {}
    void lap(const char * what, bool sync = false)
    {
        if (!on) return;
        if (sync) cudaStreamSynchronize(st);
        auto t1 = std::chrono::steady_clock::now();
        char b[96];
        snprintf(b, sizeof b, " %s=%.3f", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        log += b; t0 = t1;
    }
    ~HostTrace() { if (on) fprintf(stderr, "[lnr trace]%s\n", log.c_str()); }
};

static int apxmap_core(lnr_ctx * ctx, const lnr_index * ix, const lnr_feats * f2, const lnr_params * prm, uint32_t n_reads,
                       const u8 * d_bases, const uint64_t * h_read_off, u64 * d_out, u64 * d_out_off, uint64_t out_cap,
                       uint64_t * n_cords_total, lnr_debug_out * dbg)
{
    HostTrace tr(ctx->stream);
    ctx->anchors_host_total = 0;
    if (n_reads == 0) { if (n_cords_total) *n_cords_total = 0; return LNR_OK; }
    const float stop_ratio = prm && prm->preset == 0 ? 0.7f : 0.0f;
    const bool no_chain = prm && prm->no_chain;
    if (no_chain && ix->index_type == 2)
    {
        // -c 0 with -i 2 maps nothing in the reference: getHIndexMatchAll takes its record window from the x fields of
        // map_str / map_end (pmpfinder.cpp:1932-1933, idx_end = getCordX(map_end)) and this mode passes the bare read length
        // as map_end (:2778), so idx_end = 0, no record is accepted in either attempt and every read comes back without cords
        CK(ctx->out_off.reserve((size_t)(n_reads + STILE + 1) * sizeof(u64)));
        CK(cudaMemsetAsync(d_out_off ? (void *)d_out_off : ctx->out_off.p, 0, ((size_t)n_reads + 1) * sizeof(u64), ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        memset(ctx->counters, 0, sizeof ctx->counters);
        ctx->counters[6] = h_read_off[n_reads];
        for (uint32_t r = 0; r < n_reads; r++) ctx->counters[7] += h_read_off[r + 1] - h_read_off[r] > (u64)kMinReadLen;   // second attempts
        if (n_cords_total) *n_cords_total = 0;
        return LNR_OK;
    }
    const int ft = f2->feature_type;
    if (prm && prm->feature_type != 0 && prm->feature_type != ft) return fail(ctx, LNR_E_ARG, "lnr_params.feature_type differs from the genome features' type");
    const u32 fe_tile = ft == 1 ? (u32)(FT - 1) : (u32)FE;      // entries per feature tile
    const size_t fesz = ft == 1 ? sizeof(i16) : sizeof(F96);
    // ---- host-side layout from the read lengths
    std::vector<SeedTask> tasks(n_reads);
    std::vector<u64> foff(n_reads + 1), cbase(n_reads + 1), hoff(n_reads + 1);
    std::vector<u32> ftile(n_reads + 1);
    u64 n_samples = 0, nf_tot = 0, c_tot = 0;
    u32 n_ftiles = 0;
    for (uint32_t r = 0; r < n_reads; r++)
    {
        u64 L = h_read_off[r + 1] - h_read_off[r];
        if (L >= (1ULL << 20)) return fail(ctx, LNR_E_LIMIT, "read length must be < 2^20 (cords.h:25)");
        SeedTask & t = tasks[r];
        memset(&t, 0, sizeof t);
        t.read = r; t.str = 0; t.end = (u32)L; t.alpha = 15;
        t.n_samples = L > (u64)kMinReadLen ? (ix->index_type == 2 ? hseed_task_samples(0, (u32)L, 15) : seed_task_samples(0, (u32)L, 15)) : 0;
        t.sample0 = n_samples;
        n_samples += t.n_samples;
        u32 nf = L > (u64)kMinReadLen ? (ft == 1 ? feat32_count(L) : feat_count_read(L)) : 0;
        foff[r] = nf_tot;
        nf_tot += 2ull * nf;
        ftile[r] = n_ftiles;
        n_ftiles += 2 * ((nf + fe_tile - 1) / fe_tile);
        cbase[r] = c_tot;
        c_tot += L > (u64)kMinReadLen ? 16 + L / 4 : 0;
    }
    foff[n_reads] = nf_tot; ftile[n_reads] = n_ftiles; cbase[n_reads] = c_tot;

    const u64 total_bases = h_read_off[n_reads];
    tr.lap("layout");
    // ---- uploads
    CK(ctx->read_off.reserve((n_reads + 1) * sizeof(u64)));
    CK(ctx->tasks.reserve(n_reads * sizeof(SeedTask)));
    CK(ctx->foff.reserve((n_reads + 1) * sizeof(u64)));
    CK(ctx->ftile.reserve((n_reads + 1) * sizeof(u32)));
    CK(ctx->cords_base.reserve((n_reads + 1) * sizeof(u64)));
    CK(ctx->order.reserve((size_t)n_reads * sizeof(u32)));
    CK(ctx->feats.reserve((size_t)(nf_tot + 8) * fesz));
    CK(ctx->cords.reserve((size_t)(c_tot + 8) * sizeof(u64)));
    CK(ctx->slots.reserve((size_t)n_reads * sizeof(ReadSlot)));
    CK(ctx->ncords.reserve((size_t)(n_reads + STILE + 1) * sizeof(u32)));
    CK(ctx->out_off.reserve((size_t)(n_reads + STILE + 1) * sizeof(u64)));
    CK(ctx->misc.reserve(1024));
    CK(ctx->stage.reserve(((size_t)n_reads + 1) * (3 * sizeof(u64) + 2 * sizeof(u32)) + ((size_t)n_reads * 5 + 1024) * sizeof(SeedTask) + 1024));
    CK(upload_small(ctx, ctx->read_off.p, h_read_off, (n_reads + 1) * sizeof(u64)));
    CK(upload_small(ctx, ctx->tasks.p, tasks.data(), n_reads * sizeof(SeedTask)));
    CK(upload_small(ctx, ctx->foff.p, foff.data(), (n_reads + 1) * sizeof(u64)));
    CK(upload_small(ctx, ctx->ftile.p, ftile.data(), (n_reads + 1) * sizeof(u32)));
    CK(upload_small(ctx, ctx->cords_base.p, cbase.data(), (n_reads + 1) * sizeof(u64)));
    CK(cudaMemsetAsync(ctx->misc.p, 0, 1024, ctx->stream));

    const u64 * d_read_off = ctx->read_off.as<u64>();
    unsigned long long * d_counters = (unsigned long long *)(ctx->misc.as<u64>() + 8);
    u32 * d_queue = (u32 *)(ctx->misc.as<u64>() + 20);
    u32 * d_ntasks2 = d_queue + 1;
    u32 * d_nfail = d_queue + 2;
    tr.lap("uploads");
    // ---- read features (both strands)
    if (n_ftiles)
    {
        CK(ctx->tile_read.reserve((size_t)n_ftiles * sizeof(u32)));
        CK(ctx->feat_pairs.reserve((size_t)(n_ftiles / 2 + 1) * sizeof(FeatPair)));
        LaunchScope ls(ctx, "k_feat_reads", 2);
        if (ft == 1) k_feat_tile_reads<<<(n_reads + 255) / 256, 256, 0, ctx->stream>>>(ctx->ftile.as<u32>(), n_reads, ctx->tile_read.as<u32>());
        if (ft == 1)
            k_feat32_reads<<<std::min<u32>(n_ftiles, (u32)ctx->n_sm * 16), FT, 0, ctx->stream>>>(d_bases, d_read_off, ctx->foff.as<u64>(), ctx->ftile.as<u32>(),
                                                                                             ctx->tile_read.as<u32>(), n_ftiles, ctx->feats.as<i16>());
        else
        {
            k_feat_pairs<<<(n_reads + 255) / 256, 256, 0, ctx->stream>>>(ctx->ftile.as<u32>(), d_read_off, ctx->foff.as<u64>(), n_reads, ctx->feat_pairs.as<FeatPair>());
            k_feat_reads<<<std::min<u32>(n_ftiles / 2, (u32)ctx->n_sm * 8), FT, 0, ctx->stream>>>(d_bases, ctx->feat_pairs.as<FeatPair>(), n_ftiles, ctx->feats.as<F96>());
        }
    }
    CK(cudaGetLastError());
    // ---- primary seeding
    DevBuf & aoff = ctx->read_meta;   // reused as the per-sample anchor offset buffer
    u64 total_anchors = 0, max_task_anchors = 0;
    int rc = seeding_pass(ctx, ix, d_bases, d_read_off, ctx->tasks.as<SeedTask>(), n_reads, n_samples, aoff, &total_anchors, "k_seed_count", &max_task_anchors);
    if (rc) return rc;
    tr.lap("feat+seed(sync)");
    if (dbg && dbg->raw_anchors_off)
    {
        // per-read raw anchors without the sentinel: copy region by region (debug path, not timed)
        std::vector<u64> h_aoff(n_samples + 1);
        CK(cudaMemcpyAsync(h_aoff.data(), aoff.p, (n_samples + 1) * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        u64 o = 0;
        dbg->raw_anchors_off[0] = 0;
        for (uint32_t r = 0; r < n_reads; r++)
        {
            u64 a0 = h_aoff[tasks[r].sample0], a1 = h_aoff[tasks[r].sample0 + tasks[r].n_samples];
            u64 n = a1 - a0;
            if (dbg->raw_anchors && o + n <= dbg->raw_anchors_cap && n)
                CK(cudaMemcpyAsync(dbg->raw_anchors + o, ctx->anchorsA.as<u64>() + a0 + r + 1, n * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
            o += n;
            dbg->raw_anchors_off[r + 1] = o;
        }
        CK(cudaStreamSynchronize(ctx->stream));
    }
    // ---- workspace of the pipeline kernels
    // A warp owns one task at a time, so a batch of n reads never keeps more than n warps busy: grids, per-warp arenas and
    // histograms are sized by min(what the device holds, the batch). A 64-read block (the reference's p_calRecords pattern)
    // then costs a context ~0.3 GB of scratch instead of the 19 GB a 65 536-read batch uses.
    const int wpc = 4;
    const u64 batch_warps = ((u64)n_reads + wpc - 1) / wpc * wpc;
    auto warps_of = [&](int ctas_per_sm) { return std::min<u64>((u64)ctx->n_sm * ctas_per_sm * wpc, batch_warps); };
    auto grid_of = [&](int ctas_per_sm) { return (unsigned)(warps_of(ctas_per_sm) / wpc); };
    const int n_ctas = (int)grid_of(ctx->map_ctas_per_sm);
    const int max_cps = std::max(std::max(ctx->map_ctas_per_sm, ctx->sort_ctas_per_sm), std::max(ctx->chain_ctas_per_sm, ctx->blocks_ctas_per_sm));
    const u64 n_warps = warps_of(max_cps);    // every warp of the widest kernel owns a histogram
    // one scratch arena, divided evenly among the warps of whichever kernel is running: arena_bytes_per_warp is what a warp
    // of the map_ctas_per_sm-wide kernels gets, wider kernels (k_hits_sort needs 12 B per anchor) get proportionally less
    const size_t arena_total = (size_t)warps_of(ctx->map_ctas_per_sm) * ctx->arena_bytes_per_warp;
    auto arena_share = [&](int ctas_per_sm) { return (u64)((arena_total / (size_t)warps_of(ctas_per_sm)) & ~(size_t)255); };
    CK(ctx->bins.reserve((size_t)n_warps * kNumBins * sizeof(u32)));
    CK(ctx->arena.reserve(arena_total));
    {
        // the kernels return the histograms zeroed; only memory they have not seen yet needs the memset
        const size_t need = (size_t)n_warps * kNumBins * sizeof(u32);
        if (ctx->bins_zeroed != ctx->bins.p || ctx->bins_zeroed_cap < need)
        {
            CK(cudaMemsetAsync(ctx->bins.p, 0, need, ctx->stream));
            ctx->bins_zeroed = ctx->bins.p;
            ctx->bins_zeroed_cap = need;
        }
    }
    u64 tasks2_cap64 = 1024;   // a read emits at most L/1000 + 4 gap tasks (k_map_finish: gcap)
    for (uint32_t r = 0; r < n_reads; r++) tasks2_cap64 += (h_read_off[r + 1] - h_read_off[r]) / 1000 + 4;
    if (tasks2_cap64 > 0xfffffff0ull) return fail(ctx, LNR_E_LIMIT, "batch too large");
    u32 tasks2_cap = (u32)tasks2_cap64;
    CK(ctx->tasks2.reserve((size_t)tasks2_cap * sizeof(SeedTask)));
    if (dbg && dbg->hits_off)
    {
        u64 o = 0;
        for (uint32_t r = 0; r < n_reads; r++) { hoff[r] = o; o += (h_read_off[r + 1] - h_read_off[r]) / 8 + 64; }
        hoff[n_reads] = o;
        CK(ctx->dbg_hits.reserve((size_t)o * sizeof(u64)));
        CK(ctx->dbg_hoff.reserve((n_reads + 1) * sizeof(u64)));
        CK(ctx->dbg_nhits.reserve(n_reads * sizeof(u32)));
        CK(cudaMemsetAsync(ctx->dbg_nhits.p, 0, n_reads * sizeof(u32), ctx->stream));
        CK(cudaMemcpyAsync(ctx->dbg_hoff.p, hoff.data(), (n_reads + 1) * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (dbg && dbg->cords1_off)
    {
        CK(ctx->dbg_c1.reserve((size_t)(c_tot + 8) * sizeof(u64)));
        CK(ctx->dbg_nc1.reserve(n_reads * sizeof(u32)));
        CK(cudaMemsetAsync(ctx->dbg_nc1.p, 0, n_reads * sizeof(u32), ctx->stream));
    }
    MapArgs a;
    memset(&a, 0, sizeof a);
    a.read_off = d_read_off; a.bases = d_bases; a.n_reads = n_reads;
    a.feats = ctx->feats.as<F96>(); a.foff = ctx->foff.as<u64>();
    a.f2 = f2->d_ptrs; a.nf2 = f2->d_n;
    a.ft = ft; a.win = ft == 1 ? (u32)kWin32 : (u32)kWin;
    a.tasks = ctx->tasks.as<SeedTask>(); a.n_tasks = n_reads;
    a.aoff = aoff.as<u64>();
    a.A = ctx->anchorsA.as<u64>(); a.B = ctx->anchorsB.as<u64>();
    a.cords = ctx->cords.as<u64>(); a.cords_base = ctx->cords_base.as<u64>();
    a.slots = ctx->slots.as<ReadSlot>();
    a.tasks2 = ctx->tasks2.as<SeedTask>(); a.tasks2_cap = tasks2_cap; a.n_tasks2 = d_ntasks2;
    a.bins = ctx->bins.as<u32>(); a.arena = ctx->arena.as<u8>(); a.arena_per_warp = arena_share(ctx->map_ctas_per_sm);
    a.queue = d_queue;
    a.index_type = ix->index_type;
    a.order = ctx->order.as<u32>();
    a.stop_ratio = stop_ratio;
    a.counters = d_counters;
    if (dbg && dbg->hits_off) { a.dbg_hits = ctx->dbg_hits.as<u64>(); a.dbg_hoff = ctx->dbg_hoff.as<u64>(); a.dbg_nhits = ctx->dbg_nhits.as<u32>(); }
    if (dbg && dbg->cords1_off) { a.dbg_c1 = ctx->dbg_c1.as<u64>(); a.dbg_nc1 = ctx->dbg_nc1.as<u32>(); }
    tr.lap("seed_fill", true);
    // the most anchors (+ sentinel) a task may have to be served from a warp's share of the arena in all three section
    // kernels; the tasks above it are listed by k_order_tasks and taken by the big-arena warps of those kernels
    const u64 sec_fit_cap = std::min(arena_share(ctx->chain_ctas_per_sm), arena_share(ctx->blocks_ctas_per_sm));
    u32 fit_n = 0;
    {
        const u64 per = phase_map_scratch_bound(1) - phase_map_scratch_bound(0);       // the bound is linear in n
        const u64 n_a = sec_fit_cap > phase_map_scratch_bound(0) ? (sec_fit_cap - phase_map_scratch_bound(0)) / per : 0;
        const u64 share_sort = arena_share(ctx->sort_ctas_per_sm);
        const u64 n_b = share_sort > 4096 ? (share_sort - 4096) / 16 : 0;
        fit_n = (u32)std::min<u64>(std::min(n_a, n_b), 0x7fffffffu);
    }
    const bool heavy_lane = !no_chain && !getenv("LNR_MONOLITHIC_HITS") && !getenv("LNR_NO_HEAVY_LANE");
    CK(ctx->heavy_list.reserve((size_t)n_reads * sizeof(u32)));
    u32 * d_queue_h = d_queue + 4, * d_n_heavy = d_queue + 5;
    {
        LaunchScope ls(ctx, "k_order_tasks");
        k_order_tasks<<<1, 1024, 0, ctx->stream>>>(ctx->tasks.as<SeedTask>(), n_reads, aoff.as<u64>(), ctx->order.as<u32>(), fit_n,
                                                  heavy_lane ? ctx->heavy_list.as<u32>() : nullptr, d_n_heavy);
    }
    CK(ctx->task_nhits.reserve((size_t)std::max<u32>(n_reads, tasks2_cap) * sizeof(u32)));
    a.task_nhits = ctx->task_nhits.as<u32>();
    CK(ctx->task_state.reserve((size_t)n_reads * sizeof(u32)));
    a.task_state = ctx->task_state.as<u32>();
    CK(cudaMemsetAsync(ctx->slots.p, 0, (size_t)n_reads * sizeof(ReadSlot), ctx->stream));
    // big-arena pass: 32 warps; no task of this batch can need more than the bound for all of the batch's anchors, no read
    // more than ~200 B per cord slot in the finish stage
    u64 max_len = 0;
    for (uint32_t r = 0; r < n_reads; r++) max_len = std::max<u64>(max_len, h_read_off[r + 1] - h_read_off[r]);
    auto big_need = [&](u64 anchors) {
        u64 need = std::max<u64>(phase_map_scratch_bound((int)std::min<u64>(anchors + 2, 0x7ffffff0ull)), 256 * (16 + max_len / 4) + 65536);
        need = (need + (1u << 20) - 1) & ~(u64)((1u << 20) - 1);
        return (size_t)std::min<u64>(need, ctx->big_arena_bytes_per_warp);
    };
    size_t big_per_warp = big_need(max_task_anchors);     // a warp of the big-arena launches runs one task at a time
    CK(ctx->big_arena.reserve(32 * big_per_warp));
    CK(ctx->big_list.reserve((size_t)std::max<u32>(n_reads, tasks2_cap) * sizeof(u32)));
    a.big_arena = ctx->big_arena.as<u8>(); a.big_arena_per_warp = big_per_warp;
    a.big_list = ctx->big_list.as<u32>(); a.n_big = d_queue + 3;
    const bool want_rec = getenv("LNR_LONGEST_PROFILE") != nullptr;
    if (want_rec)
    {
        CK(ctx->warp_rec.reserve((size_t)n_warps * 48 * sizeof(u64)));
        CK(cudaMemsetAsync(ctx->warp_rec.p, 0, (size_t)n_warps * 48 * sizeof(u64), ctx->stream));
        a.warp_rec = ctx->warp_rec.as<u64>();
        a.warp_rec_stage_off = n_warps * 24;
        a.warp_rec_stage_stride = n_warps;
    }
    u32 n_tasks2 = 0;
    if (no_chain)
    {
        // -c 0: one kernel per attempt (k_map_c0) plus its big-arena launch; the second attempt is seeded at step 7 for the
        // reads the first one left too short
        if (c0_scratch_bound((int)std::min<u64>(max_task_anchors + 4, 0x7ffffff0ull)) > big_per_warp)
        {
            big_per_warp = (size_t)std::min<u64>((c0_scratch_bound((int)std::min<u64>(max_task_anchors + 4, 0x7ffffff0ull)) + (1u << 20) - 1) & ~(u64)((1u << 20) - 1),
                                                 ctx->big_arena_bytes_per_warp);
            CK(ctx->big_arena.reserve(32 * big_per_warp));
            a.big_arena = ctx->big_arena.as<u8>(); a.big_arena_per_warp = big_per_warp;
        }
        const int gdl = prm->gdl_state ? 1 : 0;
        auto run_attempt = [&](int attempt, int list_n, int best_n, const char * tag, const char * tag_big) -> int {
            CK(cudaMemsetAsync(d_queue, 0, sizeof(u32), ctx->stream));
            CK(cudaMemsetAsync(d_queue + 3, 0, sizeof(u32), ctx->stream));
            {
                LaunchScope ls(ctx, tag);
                k_map_c0<<<n_ctas, wpc * 32, 0, ctx->stream>>>(a, attempt, 0, list_n, best_n);
            }
            CK(cudaMemsetAsync(d_queue, 0, sizeof(u32), ctx->stream));
            {
                LaunchScope ls(ctx, tag_big);
                k_map_c0<<<8, 128, 0, ctx->stream>>>(a, attempt, 1, list_n, best_n);
            }
            CK(cudaGetLastError());
            return LNR_OK;
        };
        rc = run_attempt(0, gdl ? 10 : 20, gdl ? 999 : 1, "k_map_c0", "k_map_c0_big");
        if (rc) return rc;
        CK(cudaMemcpyAsync(&n_tasks2, d_ntasks2, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (n_tasks2 > tasks2_cap) return fail(ctx, LNR_E_CAPACITY, "second-attempt task buffer exhausted");
        if (n_tasks2)
        {
            std::vector<SeedTask> t2(n_tasks2);
            CK(cudaMemcpyAsync(t2.data(), ctx->tasks2.p, n_tasks2 * sizeof(SeedTask), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            std::sort(t2.begin(), t2.end(), [](const SeedTask & x, const SeedTask & y) { return x.read < y.read; });   // deterministic layout
            u64 ns2 = 0;
            for (u32 i = 0; i < n_tasks2; i++) { t2[i].sample0 = ns2; ns2 += t2[i].n_samples; }
            CK(upload_small(ctx, ctx->tasks2.p, t2.data(), n_tasks2 * sizeof(SeedTask)));
            u64 total2 = 0, max2 = 0;
            rc = seeding_pass(ctx, ix, d_bases, d_read_off, ctx->tasks2.as<SeedTask>(), n_tasks2, ns2, aoff, &total2, "k_seed_count_c0_retry", &max2);
            if (rc) return rc;
            const u64 need2 = (c0_scratch_bound((int)std::min<u64>(max2 + 4, 0x7ffffff0ull)) + (1u << 20) - 1) & ~(u64)((1u << 20) - 1);
            if (std::min<u64>(need2, ctx->big_arena_bytes_per_warp) > big_per_warp)
            {
                big_per_warp = (size_t)std::min<u64>(need2, ctx->big_arena_bytes_per_warp);
                CK(ctx->big_arena.reserve(32 * big_per_warp));
            }
            a.big_arena = ctx->big_arena.as<u8>(); a.big_arena_per_warp = big_per_warp;
            a.tasks = ctx->tasks2.as<SeedTask>(); a.n_tasks = n_tasks2;
            a.aoff = aoff.as<u64>();
            a.A = ctx->anchorsA.as<u64>(); a.B = ctx->anchorsB.as<u64>();
            rc = run_attempt(1, 20, 1, "k_map_c0_retry", "k_map_c0_retry_big");
            if (rc) return rc;
        }
    }
    else
    {
    if (getenv("LNR_MONOLITHIC_HITS"))
    {
        LaunchScope ls(ctx, "k_map_hits");
        k_map_hits<<<n_ctas, wpc * 32, 0, ctx->stream>>>(a, 0, 0);
    }
    else
    {
        // the three sections of the hit stage, one kernel each (see k_hits_sort)
        MapArgs as = a;
        as.arena_per_warp = arena_share(ctx->sort_ctas_per_sm);
        const u64 cap_chain = arena_share(ctx->chain_ctas_per_sm), cap_blocks = arena_share(ctx->blocks_ctas_per_sm);
        as.fit_cap = std::min(cap_chain, cap_blocks);   // a task must fit every section's arena (this section: 12 B per anchor)
        if (heavy_lane) { as.fit_n = fit_n; as.heavy_list = ctx->heavy_list.as<u32>(); as.n_heavy = d_n_heavy; as.queue_h = d_queue_h; }
        {
            LaunchScope ls(ctx, "k_hits_sort");
            k_hits_sort<<<grid_of(ctx->sort_ctas_per_sm), 128, 0, ctx->stream>>>(as);
        }
        CK(cudaMemsetAsync(d_queue, 0, sizeof(u32), ctx->stream));
        CK(cudaMemsetAsync(d_queue_h, 0, sizeof(u32), ctx->stream));
        as.arena_per_warp = cap_chain;
        CK(ctx->order2.reserve((size_t)n_reads * sizeof(u32)));
        as.order = ctx->order2.as<u32>();
        {
            LaunchScope ls(ctx, "k_order_by_key");
            k_order_by_key<<<1, 1024, 0, ctx->stream>>>(a.task_state, n_reads, ctx->order2.as<u32>());
        }
        {
            LaunchScope ls(ctx, "k_hits_chain");
            k_hits_chain<<<grid_of(ctx->chain_ctas_per_sm), 128, 0, ctx->stream>>>(as);
        }
        CK(cudaMemsetAsync(d_queue, 0, sizeof(u32), ctx->stream));
        CK(cudaMemsetAsync(d_queue_h, 0, sizeof(u32), ctx->stream));
        as.arena_per_warp = cap_blocks;
        {
            LaunchScope ls(ctx, "k_order_by_key");
            k_order_by_key<<<1, 1024, 0, ctx->stream>>>(a.task_state, n_reads, ctx->order2.as<u32>());
        }
        {
            LaunchScope ls(ctx, "k_hits_blocks");
            k_hits_blocks<<<grid_of(ctx->blocks_ctas_per_sm), 128, 0, ctx->stream>>>(as);
        }
    }
    CK(cudaGetLastError());
    tr.lap("hits", true);
    CK(cudaMemsetAsync(d_queue, 0, sizeof(u32), ctx->stream));
    {
        LaunchScope ls(ctx, "k_map_hits_big");
        k_map_hits<<<8, 128, 0, ctx->stream>>>(a, 0, 1);
    }
    tr.lap("hits_big", true);
    {
        LaunchScope ls(ctx, "k_map_extend");
        k_map_extend<<<(u32)(((u64)n_reads * ctx->extend_group + 127) / 128), 128, 0, ctx->stream>>>(a, ctx->order.as<u32>(), n_reads, 0, ctx->extend_group);
    }
    tr.lap("extend", true);
    CK(cudaMemsetAsync(d_queue, 0, sizeof(u32), ctx->stream));
    CK(cudaMemsetAsync(d_queue + 3, 0, sizeof(u32), ctx->stream));
    {
        LaunchScope ls(ctx, "k_map_finish");
        k_map_finish<<<n_ctas, wpc * 32, 0, ctx->stream>>>(a, ctx->order.as<u32>(), n_reads, 0, 0);
    }
    CK(cudaMemsetAsync(d_queue, 0, sizeof(u32), ctx->stream));
    {
        LaunchScope ls(ctx, "k_map_finish_big");
        k_map_finish<<<8, 128, 0, ctx->stream>>>(a, ctx->order.as<u32>(), n_reads, 0, 1);
    }
    CK(cudaGetLastError());
    tr.lap("launch_map");
    // ---- re-map pass
    CK(cudaMemcpyAsync(&n_tasks2, d_ntasks2, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    tr.lap("map(sync)");
    if (n_tasks2 > tasks2_cap) return fail(ctx, LNR_E_CAPACITY, "re-map task buffer exhausted");
    if (n_tasks2)
    {
        std::vector<SeedTask> t2(n_tasks2);
        std::vector<ReadSlot> slots(n_reads);
        CK(cudaMemcpyAsync(t2.data(), ctx->tasks2.p, n_tasks2 * sizeof(SeedTask), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(slots.data(), ctx->slots.p, n_reads * sizeof(ReadSlot), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        u64 ns2 = 0;
        for (u32 i = 0; i < n_tasks2; i++) { t2[i].sample0 = ns2; ns2 += t2[i].n_samples; }
        std::vector<u32> remap_reads;
        for (uint32_t r = 0; r < n_reads; r++) if (slots[r].status == 1) remap_reads.push_back(r);
        CK(upload_small(ctx, ctx->tasks2.p, t2.data(), n_tasks2 * sizeof(SeedTask)));
        CK(ctx->remap_list.reserve(std::max<size_t>(remap_reads.size(), 1) * sizeof(u32)));
        u32 * d_remap = ctx->remap_list.as<u32>();
        CK(upload_small(ctx, d_remap, remap_reads.data(), remap_reads.size() * sizeof(u32)));
        u64 total2 = 0, max2 = 0;
        rc = seeding_pass(ctx, ix, d_bases, d_read_off, ctx->tasks2.as<SeedTask>(), n_tasks2, ns2, aoff, &total2, "k_seed_count_remap", &max2);
        if (rc) return rc;
        if (big_need(max2) > big_per_warp)
        {
            big_per_warp = big_need(max2);
            CK(ctx->big_arena.reserve(32 * big_per_warp));
            a.big_arena = ctx->big_arena.as<u8>(); a.big_arena_per_warp = big_per_warp;
        }
        a.tasks = ctx->tasks2.as<SeedTask>(); a.n_tasks = n_tasks2;
        a.aoff = aoff.as<u64>();
        a.A = ctx->anchorsA.as<u64>(); a.B = ctx->anchorsB.as<u64>();
        a.dbg_hits = nullptr; a.dbg_c1 = nullptr; a.dbg_nhits = nullptr; a.dbg_nc1 = nullptr;
        CK(cudaMemsetAsync(d_queue, 0, sizeof(u32), ctx->stream));
        CK(cudaMemsetAsync(d_queue + 3, 0, sizeof(u32), ctx->stream));
        {
            LaunchScope ls(ctx, "k_map_hits_remap");
            k_map_hits<<<n_ctas, wpc * 32, 0, ctx->stream>>>(a, 1, 0);
        }
        CK(cudaMemsetAsync(d_queue, 0, sizeof(u32), ctx->stream));
        {
            LaunchScope ls(ctx, "k_map_hits_remap_big");
            k_map_hits<<<8, 128, 0, ctx->stream>>>(a, 1, 1);
        }
        {
            LaunchScope ls(ctx, "k_map_extend_remap");
            k_map_extend<<<(u32)(((u64)remap_reads.size() * ctx->extend_group + 127) / 128), 128, 0, ctx->stream>>>(a, d_remap, (u32)remap_reads.size(), 1, ctx->extend_group);
        }
        CK(cudaMemsetAsync(d_queue, 0, sizeof(u32), ctx->stream));
        CK(cudaMemsetAsync(d_queue + 3, 0, sizeof(u32), ctx->stream));
        {
            LaunchScope ls(ctx, "k_map_finish_remap");
            k_map_finish<<<n_ctas, wpc * 32, 0, ctx->stream>>>(a, d_remap, (u32)remap_reads.size(), 1, 0);
        }
        CK(cudaMemsetAsync(d_queue, 0, sizeof(u32), ctx->stream));
        {
            LaunchScope ls(ctx, "k_map_finish_remap_big");
            k_map_finish<<<8, 128, 0, ctx->stream>>>(a, d_remap, (u32)remap_reads.size(), 1, 1);
        }
        CK(cudaGetLastError());
    }
    }   // f_chain = 1
    tr.lap("remap", true);
    // ---- compaction into the caller's layout
    {
        LaunchScope ls(ctx, "k_slot_counts");
        k_slot_counts<<<(n_reads + 255) / 256, 256, 0, ctx->stream>>>(ctx->slots.as<ReadSlot>(), n_reads, ctx->ncords.as<u32>(), d_nfail);
    }
    CK(cudaMemsetAsync(ctx->ncords.as<u32>() + n_reads, 0, sizeof(u32), ctx->stream));
    u64 * d_total = ctx->misc.as<u64>();
    rc = device_scan<u64>(ctx, ctx->ncords.as<u32>(), (u64)n_reads + 1, 0, d_out_off ? d_out_off : ctx->out_off.as<u64>(), d_total, "k_scan_cords");
    if (rc) return rc;
    const u64 * d_off = d_out_off ? d_out_off : ctx->out_off.as<u64>();
    {
        LaunchScope ls(ctx, "k_gather_cords");
        k_gather_cords<<<(u32)(((u64)n_reads * 32 + 255) / 256), 256, 0, ctx->stream>>>(ctx->cords.as<u64>(), ctx->cords_base.as<u64>(),
                                                                                       ctx->ncords.as<u32>(), d_off, n_reads, d_out, out_cap);
    }
    CK(cudaGetLastError());
    u64 h_misc[128];
    CK(cudaMemcpyAsync(h_misc, ctx->misc.p, sizeof h_misc, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    tr.lap("gather(sync)");
    u64 total_cords = h_misc[0];
    u32 n_fail = ((u32 *)(h_misc + 20))[2];
    ctx->counters[0] = n_samples;                 // S
    ctx->counters[1] = h_misc[8 + 1];             // H
    ctx->counters[2] = h_misc[8 + 2] + ctx->anchors_host_total;   // A (HIndex: device counter; DIndex: the scans' totals)
    ctx->counters[3] = h_misc[8 + 3];             // Hits
    ctx->counters[4] = h_misc[8 + 4];             // W
    ctx->counters[5] = total_cords;               // C
    ctx->counters[6] = total_bases;
    ctx->counters[7] = n_tasks2;
    for (int i = 0; i < 8; i++) ctx->diag[i] = h_misc[8 + 16 + i];
    for (int i = 0; i < 16; i++) ctx->stage_cycles[i] = h_misc[8 + 24 + i];
    if (want_rec)
    {
        std::vector<u64> rec((size_t)n_warps * 48);
        CK(cudaMemcpy(rec.data(), ctx->warp_rec.p, rec.size() * sizeof(u64), cudaMemcpyDeviceToHost));
        size_t best = 0;
        for (size_t wq = 0; wq < (size_t)n_warps; wq++) if (rec[wq * 24 + 12] > rec[best * 24 + 12]) best = wq;
        for (int i = 0; i < 16; i++) ctx->longest_cycles[i] = rec[best * 24 + i];
        // tail analysis of the three hit-section kernels on stderr: when warps retire, and the slowest tasks
        const char * names[3] = {"k_hits_sort", "k_hits_chain", "k_hits_blocks"};
        const int cps[3] = {ctx->sort_ctas_per_sm, ctx->chain_ctas_per_sm, ctx->blocks_ctas_per_sm};
        for (int st = 0; st < 3; st++)
        {
            const size_t nw = std::min<size_t>((size_t)ctx->n_sm * cps[st] * 4, (size_t)n_warps);
            const u64 * R = rec.data() + n_warps * 24 + (size_t)st * n_warps * 8;
            u64 t0 = ~0ULL, t1 = 0;
            std::vector<double> ends;
            for (size_t wq = 0; wq < nw; wq++) if (R[wq * 8]) { t0 = std::min(t0, R[wq * 8]); t1 = std::max(t1, R[wq * 8 + 1]); }
            for (size_t wq = 0; wq < nw; wq++) if (R[wq * 8]) ends.push_back((R[wq * 8 + 1] - t0) * 1e-6);
            if (ends.empty()) continue;
            std::sort(ends.begin(), ends.end());
            fprintf(stderr, "[lnr tail] %s span %.3f ms; warp retire ms: p10 %.2f p50 %.2f p90 %.2f p99 %.2f max %.2f\n", names[st], (t1 - t0) * 1e-6,
                    ends[ends.size() / 10], ends[ends.size() / 2], ends[ends.size() * 9 / 10], ends[ends.size() * 99 / 100], ends.back());
            std::vector<size_t> idx(nw);
            for (size_t i = 0; i < nw; i++) idx[i] = i;
            std::sort(idx.begin(), idx.end(), [&](size_t x, size_t y) { return R[x * 8 + 2] > R[y * 8 + 2]; });
            for (size_t k = 0; k < std::min<size_t>(6, nw); k++)
            {
                const u64 * r = R + idx[k] * 8;
                fprintf(stderr, "[lnr tail]   task q=%llu dur %.3f ms task %llu size %llu (warp did %llu tasks, retired %.2f ms)\n", (unsigned long long)r[5], r[2] / 1.965e6,
                        (unsigned long long)r[3], (unsigned long long)r[4], (unsigned long long)r[6], (r[1] - t0) * 1e-6);
            }
        }
    }
    if (n_cords_total) *n_cords_total = total_cords;
    if (dbg && dbg->hits_off)
    {
        std::vector<u32> nh(n_reads);
        CK(cudaMemcpy(nh.data(), ctx->dbg_nhits.p, n_reads * sizeof(u32), cudaMemcpyDeviceToHost));
        u64 o = 0;
        dbg->hits_off[0] = 0;
        for (uint32_t r = 0; r < n_reads; r++)
        {
            u64 n = std::min<u64>(nh[r], hoff[r + 1] - hoff[r]);
            if (dbg->hits && o + n <= dbg->hits_cap && n)
                CK(cudaMemcpy(dbg->hits + o, ctx->dbg_hits.as<u64>() + hoff[r], n * sizeof(u64), cudaMemcpyDeviceToHost));
            o += n;
            dbg->hits_off[r + 1] = o;
        }
    }
    if (dbg && dbg->cords1_off)
    {
        std::vector<u32> nc1(n_reads);
        CK(cudaMemcpy(nc1.data(), ctx->dbg_nc1.p, n_reads * sizeof(u32), cudaMemcpyDeviceToHost));
        u64 o = 0;
        dbg->cords1_off[0] = 0;
        for (uint32_t r = 0; r < n_reads; r++)
        {
            u64 n = nc1[r];
            if (dbg->cords1 && o + n <= dbg->cords1_cap && n)
                CK(cudaMemcpy(dbg->cords1 + o, ctx->dbg_c1.as<u64>() + cbase[r], n * sizeof(u64), cudaMemcpyDeviceToHost));
            o += n;
            dbg->cords1_off[r + 1] = o;
        }
    }
    if (n_fail) return fail(ctx, LNR_E_CAPACITY, "per-read scratch exhausted for some reads (raise arena_bytes_per_warp)");
    if (total_cords > out_cap) return fail(ctx, LNR_E_CAPACITY, "cords_capacity too small");
    return LNR_OK;
}

int lnr_apxmap_batch_device(lnr_ctx * ctx, const lnr_index * ix, const lnr_feats * f2, const lnr_params * prm, uint32_t n_reads,
                            const uint8_t * dev_bases, const uint64_t * host_read_off, uint64_t * dev_cords, uint64_t * dev_cords_off,
                            uint64_t cords_capacity, uint64_t * n_cords_total)
{
    if (!ctx || !ix || !f2 || !host_read_off || !dev_bases || !dev_cords) return LNR_E_ARG;
    cudaSetDevice(ctx->device);
    return apxmap_core(ctx, ix, f2, prm, n_reads, dev_bases, host_read_off, dev_cords, dev_cords_off, cords_capacity, n_cords_total, nullptr);
}

int lnr_apxmap_batch(lnr_ctx * ctx, const lnr_index * ix, const lnr_feats * f2, const lnr_params * prm, uint32_t n_reads,
                     const uint8_t * bases, const uint64_t * read_off, uint64_t * cords, uint64_t * cords_off, uint64_t cords_capacity,
                     lnr_debug_out * dbg)
{
    if (!ctx || !ix || !f2 || !read_off || (!bases && n_reads) || !cords || !cords_off) return LNR_E_ARG;
    cudaSetDevice(ctx->device);
    if (n_reads == 0) { cords_off[0] = 0; return LNR_OK; }
    u64 total_bases = read_off[n_reads];
    CK(ctx->bases.reserve((size_t)total_bases + 256));
    {
        std::lock_guard<std::mutex> lk(g_upload_mutex);
        CK(cudaMemcpyAsync(ctx->bases.p, bases, total_bases, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(ctx->bases.as<u8>() + total_bases, 0, 256, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    CK(ctx->out_cords.reserve((size_t)(cords_capacity + 8) * sizeof(u64)));
    u64 total = 0;
    int rc = apxmap_core(ctx, ix, f2, prm, n_reads, ctx->bases.as<u8>(), read_off, ctx->out_cords.as<u64>(), nullptr, cords_capacity, &total, dbg);
    if (rc && rc != LNR_E_CAPACITY) return rc;
    if (rc == LNR_E_CAPACITY && total > cords_capacity) return rc;
    CK(cudaMemcpyAsync(cords_off, ctx->out_off.p, (n_reads + 1) * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(cords, ctx->out_cords.p, (size_t)total * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return rc;
}

// ---- 2-bit packed reads -------------------------------------------------------------------------------------------
// Wire format of lnr_apxmap_batch_packed: base i of the batch (reads back to back) is bits 2(i&3)..2(i&3)+1 of
// packed2[i >> 2]; n_mask (optional) has bit (i & 7) of n_mask[i >> 3] set where base i is N (its 2 bits are then 0).
// One quarter (three eighths with the N bitmap) of the bytes of the 1-byte Dna5 form cross PCIe; the device expands them
// to ordinals once (16 bases per thread: one 4-byte load, one 16-byte store), every kernel downstream is unchanged.
__global__ void __launch_bounds__(256) k_unpack2(const u32 * __restrict__ packed, const u16 * __restrict__ nmask, u64 n_bases, uint4 * __restrict__ out)
{
    const u64 n16 = (n_bases + 15) / 16;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (u64)gridDim.x * blockDim.x)
    {
        const u32 w = __ldg(packed + i);
        const u32 nm = nmask ? (u32)__ldg(nmask + i) : 0u;
        u32 o[4];
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
            const u32 b = (w >> (8 * q)) & 0xffu;          // 4 bases
            u32 v = (b & 3u) | ((b & 12u) << 6) | ((b & 48u) << 12) | ((b & 192u) << 18);
            const u32 n4 = (nm >> (4 * q)) & 15u;
            if (n4) v = (v & ~(((n4 & 1u) * 0xffu) | ((n4 >> 1 & 1u) * 0xff00u) | ((n4 >> 2 & 1u) * 0xff0000u) | ((n4 >> 3 & 1u) * 0xff000000u))) |
                        ((n4 & 1u) * 4u) | ((n4 >> 1 & 1u) * 0x400u) | ((n4 >> 2 & 1u) * 0x40000u) | ((n4 >> 3 & 1u) * 0x4000000u);
            o[q] = v;
        }
        out[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

int lnr_pack_dna5(const uint8_t * dna5, uint64_t n_bases, uint8_t * packed2, uint8_t * n_mask, int * has_n)
{
    if ((!dna5 && n_bases) || !packed2) return LNR_E_ARG;
    const uint64_t nm = (n_bases + 7) / 8;
    if (n_mask) memset(n_mask, 0, (size_t)nm);
    int any = 0;
    // 8 bases per step: two 32-bit multiplies gather the low 2 bits of 4 bytes each into one byte (base i at bits 2i);
    // a group with an N (any byte > 3) takes the per-base path
    uint64_t i = 0;
    const uint64_t n8 = n_bases & ~7ull;
    for (; i < n8; i += 8)
    {
        uint64_t x;
        memcpy(&x, dna5 + i, 8);
        if ((x & 0xfcfcfcfcfcfcfcfcull) == 0)
        {
            const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
            packed2[i >> 2] = (uint8_t)((lo * 0x01041040u) >> 24);
            packed2[(i >> 2) + 1] = (uint8_t)((hi * 0x01041040u) >> 24);
            continue;
        }
        uint8_t b0 = 0, b1 = 0;
        for (int k = 0; k < 8; k++)
        {
            const unsigned c = dna5[i + k];
            if (c < 4) { if (k < 4) b0 |= (uint8_t)(c << (2 * k)); else b1 |= (uint8_t)(c << (2 * (k - 4))); }
            else { any = 1; if (n_mask) n_mask[i >> 3] |= (uint8_t)(1u << k); }
        }
        packed2[i >> 2] = b0; packed2[(i >> 2) + 1] = b1;
    }
    for (; i < n_bases; i++)
    {
        if ((i & 3) == 0) packed2[i >> 2] = 0;
        const unsigned c = dna5[i];
        if (c < 4) packed2[i >> 2] |= (uint8_t)(c << (2 * (i & 3)));
        else { any = 1; if (n_mask) n_mask[i >> 3] |= (uint8_t)(1u << (i & 7)); }
    }
    if (has_n) *has_n = any;
    if (any && !n_mask) return LNR_E_ARG;     // an N needs the bitmap
    return LNR_OK;
}

int lnr_apxmap_batch_packed(lnr_ctx * ctx, const lnr_index * ix, const lnr_feats * f2, const lnr_params * prm, uint32_t n_reads,
                            const uint8_t * packed2, const uint8_t * n_mask, const uint64_t * read_off, uint64_t * cords,
                            uint64_t * cords_off, uint64_t cords_capacity, lnr_debug_out * dbg)
{
    if (!ctx || !ix || !f2 || !read_off || (!packed2 && n_reads) || !cords || !cords_off) return LNR_E_ARG;
    cudaSetDevice(ctx->device);
    if (n_reads == 0) { cords_off[0] = 0; return LNR_OK; }
    const u64 total_bases = read_off[n_reads];
    const size_t n16 = (size_t)((total_bases + 15) / 16);
    CK(ctx->bases.reserve(n16 * 16 + 256));
    CK(ctx->packed.reserve(n16 * 4 + (n_mask ? n16 * 2 : 0) + 64));
    u8 * d_packed = ctx->packed.as<u8>();
    u8 * d_nmask = n_mask ? d_packed + n16 * 4 : nullptr;
    {
        std::lock_guard<std::mutex> lk(g_upload_mutex);
        CK(cudaMemsetAsync(d_packed, 0, n16 * 4 + (n_mask ? n16 * 2 : 0), ctx->stream));   // the tail words of the last 16-base group
        CK(cudaMemcpyAsync(d_packed, packed2, (size_t)((total_bases + 3) / 4), cudaMemcpyHostToDevice, ctx->stream));
        if (n_mask) CK(cudaMemcpyAsync(d_nmask, n_mask, (size_t)((total_bases + 7) / 8), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    {
        LaunchScope ls(ctx, "k_unpack2");
        k_unpack2<<<ctx->n_sm * 8, 256, 0, ctx->stream>>>((const u32 *)d_packed, (const u16 *)d_nmask, total_bases, (uint4 *)ctx->bases.p);
    }
    CK(cudaMemsetAsync(ctx->bases.as<u8>() + total_bases, 0, n16 * 16 + 256 - total_bases, ctx->stream));   // zero slack behind the batch
    CK(ctx->out_cords.reserve((size_t)(cords_capacity + 8) * sizeof(u64)));
    u64 total = 0;
    int rc = apxmap_core(ctx, ix, f2, prm, n_reads, ctx->bases.as<u8>(), read_off, ctx->out_cords.as<u64>(), nullptr, cords_capacity, &total, dbg);
    if (rc && rc != LNR_E_CAPACITY) return rc;
    if (rc == LNR_E_CAPACITY && total > cords_capacity) return rc;
    CK(cudaMemcpyAsync(cords_off, ctx->out_off.p, (n_reads + 1) * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(cords, ctx->out_cords.p, (size_t)total * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return rc;
}

int lnr_cords_to_records(lnr_ctx * ctx, uint32_t n_reads, const uint64_t * cords, const uint64_t * cords_off, const uint64_t * read_len,
                         const lnr_bam_parms * prm, lnr_bam_rec * recs, uint64_t rec_cap, uint64_t * rec_off, uint64_t * cigars,
                         uint64_t cigar_cap, uint64_t * cigar_off)
{
    if (!ctx || !cords_off || !read_len || !prm || !rec_off || !cigar_off || (n_reads && cords_off[n_reads] && !cords)) return LNR_E_ARG;
    static_assert(sizeof(lnr_bam_rec) == sizeof(BamRec), "lnr_bam_rec layout");
    cudaSetDevice(ctx->device);
    rec_off[0] = 0; cigar_off[0] = 0;
    if (n_reads == 0) return LNR_OK;
    const u64 n_cords = cords_off[n_reads];
    BamParms P;
    P.window = prm->window ? prm->window : 96u; P.thd_large_x = prm->thd_large_x; P.thd_di = prm->thd_di; P.thd_x = prm->thd_x;
    if (P.thd_di <= 0 || P.thd_x < 0) return fail(ctx, LNR_E_ARG, "thd_di must be positive");
    // workspace: inputs, the two count arrays and their scans
    size_t off_t = 0;
    auto carve = [&](size_t bytes) { size_t o = off_t; off_t += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_cords = carve((size_t)(n_cords + 1) * 8), o_coff = carve((size_t)(n_reads + 1) * 8), o_len = carve((size_t)n_reads * 8);
    const size_t o_nrec = carve((size_t)(n_reads + STILE + 1) * 4), o_ncig = carve((size_t)(n_reads + STILE + 1) * 4);
    const size_t o_roff = carve((size_t)(n_reads + STILE + 1) * 8), o_goff = carve((size_t)(n_reads + STILE + 1) * 8), o_tot = carve(64);
    CK(ctx->out_cords.reserve(off_t));
    u8 * W = ctx->out_cords.as<u8>();
    u64 * d_cords = (u64 *)(W + o_cords); u64 * d_coff = (u64 *)(W + o_coff); u64 * d_len = (u64 *)(W + o_len);
    u32 * d_nrec = (u32 *)(W + o_nrec); u32 * d_ncig = (u32 *)(W + o_ncig);
    u64 * d_roff = (u64 *)(W + o_roff); u64 * d_goff = (u64 *)(W + o_goff); u64 * d_tot = (u64 *)(W + o_tot);
    if (n_cords) CK(cudaMemcpyAsync(d_cords, cords, (size_t)n_cords * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_coff, cords_off, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_len, read_len, (size_t)n_reads * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(d_nrec + n_reads, 0, 4, ctx->stream));
    CK(cudaMemsetAsync(d_ncig + n_reads, 0, 4, ctx->stream));
    {
        LaunchScope ls(ctx, "k_bam_count");
        k_bam_records<false><<<(n_reads + 127) / 128, 128, 0, ctx->stream>>>(d_cords, d_coff, d_len, n_reads, P, d_nrec, d_ncig, nullptr, nullptr, nullptr, nullptr);
    }
    int rc = device_scan<u64>(ctx, d_nrec, (u64)n_reads + 1, 0, d_roff, d_tot, "k_scan_bam");
    if (rc) return rc;
    rc = device_scan<u64>(ctx, d_ncig, (u64)n_reads + 1, 0, d_goff, d_tot + 1, "k_scan_bam");
    if (rc) return rc;
    u64 tot[2];
    CK(cudaMemcpyAsync(tot, d_tot, sizeof tot, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(rec_off, d_roff, (size_t)(n_reads + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(cigar_off, d_goff, (size_t)(n_reads + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (tot[0] > rec_cap || tot[1] > cigar_cap || (tot[0] && !recs) || (tot[1] && !cigars))
        return fail(ctx, LNR_E_CAPACITY, "record / cigar buffer too small (needed sizes are in rec_off[n_reads] and cigar_off[n_reads])");
    if (tot[0])
    {
        CK(ctx->cords.reserve((size_t)tot[0] * sizeof(BamRec) + (size_t)(tot[1] + 1) * 8 + 512));
        BamRec * d_recs = ctx->cords.as<BamRec>();
        u64 * d_cigs = (u64 *)(ctx->cords.as<u8>() + (((size_t)tot[0] * sizeof(BamRec) + 255) & ~(size_t)255));
        {
            LaunchScope ls(ctx, "k_bam_fill");
            k_bam_records<true><<<(n_reads + 127) / 128, 128, 0, ctx->stream>>>(d_cords, d_coff, d_len, n_reads, P, nullptr, nullptr, d_roff, d_goff, d_recs, d_cigs);
        }
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(recs, d_recs, (size_t)tot[0] * sizeof(BamRec), cudaMemcpyDeviceToHost, ctx->stream));
        if (tot[1]) CK(cudaMemcpyAsync(cigars, d_cigs, (size_t)tot[1] * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return LNR_OK;
}

int lnr_last_batch_counters(lnr_ctx * ctx, uint64_t counters[8])
{
    if (!ctx || !counters) return LNR_E_ARG;
    for (int i = 0; i < 8; i++) counters[i] = ctx->counters[i];
    return LNR_OK;
}

int lnr_last_batch_diag(lnr_ctx * ctx, uint64_t diag[8])
{
    if (!ctx || !diag) return LNR_E_ARG;
    for (int i = 0; i < 8; i++) diag[i] = ctx->diag[i];
    return LNR_OK;
}

int lnr_last_batch_stage_cycles(lnr_ctx * ctx, uint64_t cycles[16])
{
    if (!ctx || !cycles) return LNR_E_ARG;
    for (int i = 0; i < 16; i++) cycles[i] = ctx->stage_cycles[i];
    if (getenv("LNR_LONGEST_PROFILE")) for (int i = 0; i < 16; i++) cycles[i] = ctx->longest_cycles[i];
    return LNR_OK;
}

// ---- read ingest (lnr_ingest.cuh) ---------------------------------------------------------------------------------
int lnr_reads_parse_device(lnr_ctx * ctx, const char * dev_text, uint64_t n_bytes, int first_byte, int cut_id_at_space, lnr_reads ** out)
{
    if (!ctx || !out || (!dev_text && n_bytes)) return LNR_E_ARG;
    cudaSetDevice(ctx->device);
    if (n_bytes == 0)
    {
        lnr_reads * R = new lnr_reads();
        R->ctx = ctx;
        R->h_off.assign(1, 0);
        *out = R;
        return LNR_OK;
    }
    return reads_parse_device(ctx, (const u8 *)dev_text, n_bytes, first_byte, cut_id_at_space, out);
}
int lnr_reads_parse(lnr_ctx * ctx, const char * text, uint64_t n_bytes, int cut_id_at_space, lnr_reads ** out)
{
    if (!ctx || !out || (!text && n_bytes)) return LNR_E_ARG;
    cudaSetDevice(ctx->device);
    if (n_bytes == 0) return lnr_reads_parse_device(ctx, nullptr, 0, 0, cut_id_at_space, out);
    CK(ctx->bases.reserve((size_t)n_bytes + 256));   // the batch's base buffer doubles as the text staging area
    {
        std::lock_guard<std::mutex> lk(g_upload_mutex);
        CK(cudaMemcpyAsync(ctx->bases.p, text, n_bytes, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return reads_parse_device(ctx, ctx->bases.as<u8>(), n_bytes, (unsigned char)text[0], cut_id_at_space, out);
}
int lnr_reads_info(const lnr_reads * R, uint64_t * n_reads, uint64_t * total_bases)
{
    if (!R) return LNR_E_ARG;
    if (n_reads) *n_reads = R->n_reads;
    if (total_bases) *total_bases = R->total_bases;
    return LNR_OK;
}
int lnr_reads_download(const lnr_reads * R, uint8_t * bases, uint64_t * read_off, uint64_t * id_off, uint32_t * id_len)
{
    if (!R) return LNR_E_ARG;
    lnr_ctx * ctx = R->ctx;
    cudaSetDevice(ctx->device);
    if (read_off && R->n_reads == 0) read_off[0] = 0;
    if (R->n_reads == 0) return LNR_OK;
    if (bases && R->total_bases) CK(cudaMemcpyAsync(bases, R->d_bases, R->total_bases, cudaMemcpyDeviceToHost, ctx->stream));
    if (read_off) CK(cudaMemcpyAsync(read_off, R->d_off, (R->n_reads + 1) * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    if (id_off) CK(cudaMemcpyAsync(id_off, R->d_id_off, R->n_reads * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    if (id_len) CK(cudaMemcpyAsync(id_len, R->d_id_len, R->n_reads * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return LNR_OK;
}
int lnr_reads_device(const lnr_reads * R, const uint8_t ** dev_bases, const uint64_t ** dev_read_off)
{
    if (!R) return LNR_E_ARG;
    if (dev_bases) *dev_bases = R->d_bases;
    if (dev_read_off) *dev_read_off = R->d_off;
    return LNR_OK;
}
int lnr_apxmap_reads(lnr_ctx * ctx, const lnr_index * ix, const lnr_feats * f2, const lnr_params * prm, const lnr_reads * R, uint32_t first,
                     uint32_t n_reads, uint64_t * cords, uint64_t * cords_off, uint64_t cords_capacity, lnr_debug_out * dbg)
{
    if (!ctx || !ix || !f2 || !R || !cords || !cords_off || (u64)first + n_reads > R->n_reads) return LNR_E_ARG;
    cudaSetDevice(ctx->device);
    if (n_reads == 0) { cords_off[0] = 0; return LNR_OK; }
    std::vector<u64> ro(n_reads + 1);
    const u64 o0 = R->h_off[first];
    for (uint32_t i = 0; i <= n_reads; i++) ro[i] = R->h_off[first + i] - o0;
    CK(ctx->out_cords.reserve((size_t)(cords_capacity + 8) * sizeof(u64)));
    u64 total = 0;
    int rc = apxmap_core(ctx, ix, f2, prm, n_reads, R->d_bases + o0, ro.data(), ctx->out_cords.as<u64>(), nullptr, cords_capacity, &total, dbg);
    if (rc && rc != LNR_E_CAPACITY) return rc;
    if (rc == LNR_E_CAPACITY && total > cords_capacity) return rc;
    CK(cudaMemcpyAsync(cords_off, ctx->out_off.p, (n_reads + 1) * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(cords, ctx->out_cords.p, (size_t)total * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return rc;
}
void lnr_reads_destroy(lnr_reads * R)
{
    if (!R) return;
    cudaSetDevice(R->ctx->device);
    if (R->d_block)
    {
        lnr_ctx * ctx = R->ctx;
        if (R->block_bytes > ctx->reads_cache_bytes)
        {
            if (ctx->reads_cache) cudaFree(ctx->reads_cache);
            ctx->reads_cache = R->d_block; ctx->reads_cache_bytes = R->block_bytes;
        }
        else cudaFree(R->d_block);
    }
    delete R;
}

int lnr_selftest_sort(lnr_ctx * ctx, uint64_t * records, uint32_t n)
{
    if (!ctx || (!records && n)) return LNR_E_ARG;
    if (n == 0) return LNR_OK;
    cudaSetDevice(ctx->device);
    u64 * d = nullptr;
    CK(cudaMalloc(&d, (size_t)(4 * (size_t)n + 8) * sizeof(u64)));
    cudaError_t e = cudaMemcpyAsync(d, records, (size_t)n * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
    {
        k_selftest_sort<<<1, 32, 0, ctx->stream>>>(d, d + n, d + 2 * (size_t)n, (int)n, d + 3 * (size_t)n);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(records, d + 3 * (size_t)n, (size_t)n * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, LNR_E_CUDA, cudaGetErrorString(e));
    return LNR_OK;
}

int lnr_read_features(lnr_ctx * ctx, const uint8_t * dna5, uint64_t len, int feature_type, void * dst_fwd, void * dst_rev,
                      uint64_t cap_entries, uint64_t * n_entries)
{
    if (!ctx || !dna5) return LNR_E_ARG;
    if (feature_type != 2 && feature_type != 1) return fail(ctx, LNR_E_UNSUPPORTED, "feature_type must be 1 or 2");
    cudaSetDevice(ctx->device);
    const size_t esz = feature_type == 1 ? sizeof(i16) : sizeof(F96);
    const u32 fe_tile = feature_type == 1 ? (u32)(FT - 1) : (u32)FE;
    u32 nf = feature_type == 1 ? feat32_count(len) : feat_count_read(len);
    if (n_entries) *n_entries = nf;
    if (!nf || (!dst_fwd && !dst_rev)) return LNR_OK;   // count query
    if (cap_entries < nf) return fail(ctx, LNR_E_CAPACITY, "feature buffer too small");
    CK(ctx->bases.reserve((size_t)len + 256));
    CK(ctx->feats.reserve((size_t)(2 * nf + 8) * sizeof(F96)));
    CK(ctx->read_off.reserve(2 * sizeof(u64)));
    CK(ctx->foff.reserve(2 * sizeof(u64)));
    CK(ctx->ftile.reserve(2 * sizeof(u32)));
    u64 ro[2] = {0, len}, fo[2] = {0, 2ull * nf};
    u32 ft[2] = {0, 2 * ((nf + fe_tile - 1) / fe_tile)};
    CK(cudaMemcpyAsync(ctx->bases.p, dna5, len, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->read_off.p, ro, sizeof ro, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->foff.p, fo, sizeof fo, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->ftile.p, ft, sizeof ft, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->tile_read.reserve((size_t)(ft[1] + 1) * sizeof(u32)));
    CK(cudaMemsetAsync(ctx->tile_read.p, 0, (size_t)(ft[1] + 1) * sizeof(u32), ctx->stream));   // one read: every tile is read 0
    if (ft[1] && feature_type == 1)
        k_feat32_reads<<<ft[1], FT, 0, ctx->stream>>>(ctx->bases.as<u8>(), ctx->read_off.as<u64>(), ctx->foff.as<u64>(), ctx->ftile.as<u32>(),
                                                  ctx->tile_read.as<u32>(), ft[1], ctx->feats.as<i16>());
    else if (ft[1])
    {
        CK(ctx->feat_pairs.reserve((size_t)(ft[1] / 2 + 1) * sizeof(FeatPair)));
        k_feat_pairs<<<1, 32, 0, ctx->stream>>>(ctx->ftile.as<u32>(), ctx->read_off.as<u64>(), ctx->foff.as<u64>(), 1u, ctx->feat_pairs.as<FeatPair>());
        k_feat_reads<<<ft[1] / 2, FT, 0, ctx->stream>>>(ctx->bases.as<u8>(), ctx->feat_pairs.as<FeatPair>(), ft[1], ctx->feats.as<F96>());
    }
    CK(cudaGetLastError());
    if (dst_fwd) CK(cudaMemcpyAsync(dst_fwd, ctx->feats.p, (size_t)nf * esz, cudaMemcpyDeviceToHost, ctx->stream));
    if (dst_rev) CK(cudaMemcpyAsync(dst_rev, ctx->feats.as<u8>() + (size_t)nf * esz, (size_t)nf * esz, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return LNR_OK;
}

}  // extern "C"
