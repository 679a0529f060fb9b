// Warp-cooperative primitives used by the per-read pipeline (one warp owns one read).
// In the CUDA build they map to shuffle / ballot / match intrinsics; in the host build
// (tests/host_emu, plain g++) the "warp" has a single lane and every collective is the identity,
// so the same pipeline source can be checked against the oracle on a machine without a GPU.
#pragma once
#include "lnr_defs.h"

namespace lnr {

#ifdef __CUDACC__
#define LNR_PIPE __device__ __noinline__
#define LNR_PIPE_INL __device__ __forceinline__
#else
#define LNR_PIPE inline
#define LNR_PIPE_INL inline
#endif

struct Warp
{
    int lane;        // 0..nl-1
    // lanes that cooperate on one read: the hardware warp on the device, one lane in the host build. A compile-time
    // constant on purpose: the pipeline strides, unrolls and divides by it in every inner loop.
#ifdef __CUDACC__
    static constexpr int nl = 32;
#else
    static constexpr int nl = 1;
#endif
    unsigned mask;   // member mask (device only)
};

#ifdef __CUDACC__
static const unsigned kFull = 0xffffffffu;
// a Warp with nl == 1 is a single thread running the warp-uniform code on its own (thread-per-read kernels)
LNR_PIPE_INL void wsync(const Warp & w) { if (w.nl != 1) __syncwarp(w.mask); }
LNR_PIPE_INL u32 wballot(const Warp &, bool p) { return __ballot_sync(kFull, p); }
template <class T> LNR_PIPE_INL T wbcast(const Warp &, T v, int src) { return __shfl_sync(kFull, v, src); }
LNR_PIPE_INL u64 wbcast64(const Warp &, u64 v, int src)
{
    u32 lo = __shfl_sync(kFull, (u32)v, src), hi = __shfl_sync(kFull, (u32)(v >> 32), src);
    return ((u64)hi << 32) | lo;
}
LNR_PIPE_INL int wsum(const Warp & w, int v)
{
    return __reduce_add_sync(kFull, v);
}
LNR_PIPE_INL u64 wor64(const Warp &, u64 v)
{
    for (int o = 16; o; o >>= 1)
    {
        u32 lo = __shfl_xor_sync(kFull, (u32)v, o), hi = __shfl_xor_sync(kFull, (u32)(v >> 32), o);
        v |= ((u64)hi << 32) | lo;
    }
    return v;
}
LNR_PIPE_INL i64 wmax_i64(const Warp &, i64 v)
{
    for (int o = 16; o; o >>= 1)
    {
        u32 lo = __shfl_xor_sync(kFull, (u32)(u64)v, o), hi = __shfl_xor_sync(kFull, (u32)((u64)v >> 32), o);
        i64 t = (i64)(((u64)hi << 32) | lo);
        v = t > v ? t : v;
    }
    return v;
}
// exclusive prefix sum over lanes; total returned through `total`
LNR_PIPE_INL int wscan_excl(const Warp & w, int v, int & total)
{
    int s = v;
    for (int o = 1; o < 32; o <<= 1)
    {
        int t = __shfl_up_sync(kFull, s, o);
        if (w.lane >= o) s += t;
    }
    total = __shfl_sync(kFull, s, 31);
    return s - v;
}
LNR_PIPE_INL u32 wmax_u32(const Warp &, u32 v) { return __reduce_max_sync(kFull, v); }
LNR_PIPE_INL i32 wmax_i32(const Warp &, i32 v) { return __reduce_max_sync(kFull, v); }
LNR_PIPE_INL void wor_flag64(u64 * p, u64 f) { atomicOr((unsigned long long *)p, (unsigned long long)f); }
LNR_PIPE_INL u32 wmatch(const Warp &, u32 key) { return __match_any_sync(kFull, key); }
LNR_PIPE_INL u32 wshift_up32(const Warp &, u32 v) { return __shfl_up_sync(kFull, v, 1); }   // lane l gets lane l-1 (lane 0 keeps its own)
LNR_PIPE_INL u64 wshift_up64(const Warp &, u64 v)
{
    u32 lo = __shfl_up_sync(kFull, (u32)v, 1), hi = __shfl_up_sync(kFull, (u32)(v >> 32), 1);
    return ((u64)hi << 32) | lo;
}
LNR_PIPE_INL int popc_below(const Warp & w, u32 mask) { return __popc(mask & ((1u << w.lane) - 1)); }
LNR_PIPE_INL int popc32(u32 m) { return __popc(m); }
LNR_PIPE_INL int ffs32(u32 m) { return __ffs((int)m) - 1; }
LNR_PIPE_INL int hibit32(u32 m) { return 31 - __clz((int)m); }
#else
LNR_PIPE_INL void wsync(const Warp &) {}
LNR_PIPE_INL u32 wballot(const Warp &, bool p) { return p ? 1u : 0u; }
template <class T> LNR_PIPE_INL T wbcast(const Warp &, T v, int) { return v; }
LNR_PIPE_INL u64 wbcast64(const Warp &, u64 v, int) { return v; }
LNR_PIPE_INL int wsum(const Warp &, int v) { return v; }
LNR_PIPE_INL u64 wor64(const Warp &, u64 v) { return v; }
LNR_PIPE_INL i64 wmax_i64(const Warp &, i64 v) { return v; }
LNR_PIPE_INL int wscan_excl(const Warp &, int v, int & total) { total = v; return 0; }
LNR_PIPE_INL u32 wmax_u32(const Warp &, u32 v) { return v; }
LNR_PIPE_INL i32 wmax_i32(const Warp &, i32 v) { return v; }
LNR_PIPE_INL void wor_flag64(u64 * p, u64 f) { *p |= f; }
LNR_PIPE_INL u32 wmatch(const Warp &, u32) { return 1u; }
LNR_PIPE_INL u32 wshift_up32(const Warp &, u32 v) { return v; }
LNR_PIPE_INL u64 wshift_up64(const Warp &, u64 v) { return v; }
LNR_PIPE_INL int popc_below(const Warp &, u32) { return 0; }
LNR_PIPE_INL int popc32(u32 m) { return __builtin_popcount(m); }
LNR_PIPE_INL int ffs32(u32 m) { return __builtin_ffs((int)m) - 1; }
LNR_PIPE_INL int hibit32(u32 m) { return 31 - __builtin_clz(m); }
#endif

// Bump allocator over the warp's private scratch region (reset for every read). Allocation failure is
// sticky and reported per read; the pipeline never writes past a failed allocation.
struct Arena
{
    u8 * base;
    u64 cap, off;
    int failed;
};
LNR_PIPE_INL void arena_reset(Arena & a) { a.off = 0; a.failed = 0; }
template <class T> LNR_PIPE_INL T * arena_alloc(Arena & a, u64 n)
{
    u64 bytes = (n * sizeof(T) + 15) & ~15ULL;
    if (a.off + bytes > a.cap) { a.failed = 1; return (T *)0; }
    T * p = (T *)(a.base + a.off);
    a.off += bytes;
    return p;
}

}  // namespace lnr
