// Device-wide stable LSD radix sort of (u64 key, u32 aux) records, 8 bits per pass, for the HIndex build.
// One pass = k_rs_hist (per-tile digit histograms) -> device scan over (digit, tile) -> k_rs_scatter (stable local
// ranks: each warp ranks its 256-element slice 32 at a time with match_any, warp bases come from a per-tile prefix).
// The digit of a record is taken either from the key (optionally complemented = descending order) or from aux.
#pragma once

static const int RS_T = 256;                 // threads per CTA
static const int RS_IPT = 8;                 // items per thread
static const int RS_TILE = RS_T * RS_IPT;    // 2048 records per CTA

struct RsDigit { int from_aux; int shift; u64 xor_mask; };   // digit = ((from_aux ? aux : key ^ xor_mask) >> shift) & 255
__device__ __forceinline__ u32 rs_digit(const RsDigit & d, u64 key, u32 aux)
{
    return d.from_aux ? ((aux >> d.shift) & 255u) : (u32)(((key ^ d.xor_mask) >> d.shift) & 255ull);
}

// hist[digit * n_tiles + tile]
__global__ void __launch_bounds__(RS_T) k_rs_hist(const u64 * __restrict__ key, const u32 * __restrict__ aux, u64 n, RsDigit dg,
                                                  u32 * __restrict__ hist, u32 n_tiles)
{
    __shared__ u32 s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    u64 base = (u64)blockIdx.x * RS_TILE;
#pragma unroll
    for (int i = 0; i < RS_IPT; i++)
    {
        u64 idx = base + (u64)i * RS_T + threadIdx.x;
        if (idx < n) atomicAdd(&s_h[rs_digit(dg, key[idx], aux ? aux[idx] : 0u)], 1u);
    }
    __syncthreads();
    hist[(u64)threadIdx.x * n_tiles + blockIdx.x] = s_h[threadIdx.x];
}

// offs = exclusive scan of hist (same layout). Warp wp of the CTA owns records [base + wp*256, +256) in index order.
__global__ void __launch_bounds__(RS_T) k_rs_scatter(const u64 * __restrict__ key, const u32 * __restrict__ aux, u64 n, RsDigit dg,
                                                     const u64 * __restrict__ offs, u32 n_tiles, u64 * __restrict__ key_out,
                                                     u32 * __restrict__ aux_out)
{
    __shared__ u32 s_cnt[RS_T / 32][256];    // per-warp digit counts, then running positions
    __shared__ u64 s_base[256];
    const unsigned lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (RS_T / 32) * 256; i += RS_T) (&s_cnt[0][0])[i] = 0;
    s_base[threadIdx.x] = offs[(u64)threadIdx.x * n_tiles + blockIdx.x];
    __syncthreads();
    const u64 wbase = (u64)blockIdx.x * RS_TILE + (u64)wp * (RS_TILE / (RS_T / 32));
    u64 k[RS_IPT]; u32 a[RS_IPT]; u32 d[RS_IPT];
#pragma unroll
    for (int i = 0; i < RS_IPT; i++)
    {
        u64 idx = wbase + (u64)i * 32 + lane;
        bool v = idx < n;
        k[i] = v ? key[idx] : 0;
        a[i] = (v && aux) ? aux[idx] : 0u;
        d[i] = v ? rs_digit(dg, k[i], a[i]) : 0xffffffffu;
        if (v) atomicAdd(&s_cnt[wp][d[i]], 1u);
    }
    __syncthreads();
    // per digit: exclusive prefix over the warps -> each warp's first position for that digit
    {
        u32 dgt = threadIdx.x, run = 0;
        for (int w2 = 0; w2 < RS_T / 32; w2++) { u32 c = s_cnt[w2][dgt]; s_cnt[w2][dgt] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_IPT; i++)
    {
        bool v = d[i] != 0xffffffffu;
        u32 peers = __match_any_sync(0xffffffffu, v ? d[i] : 0x1000u + lane);
        u32 rank = __popc(peers & ((1u << lane) - 1));
        u32 pos = v ? s_cnt[wp][d[i]] : 0;
        __syncwarp();
        if (v)
        {
            u64 o = s_base[d[i]] + pos + rank;
            key_out[o] = k[i];
            if (aux_out) aux_out[o] = a[i];
            if (rank == (u32)__popc(peers) - 1) s_cnt[wp][d[i]] = pos + rank + 1;
        }
        __syncwarp();
    }
}

// OR / AND of all keys and aux values: tells which digits vary at all (constant digits are skipped)
__global__ void __launch_bounds__(256) k_rs_masks(const u64 * __restrict__ key, const u32 * __restrict__ aux, u64 n, u64 * __restrict__ out4)
{
    u64 ko = 0, ka = ~0ull, ao = 0, aa = ~0ull;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
    {
        u64 kk = key[i]; ko |= kk; ka &= kk;
        if (aux) { u64 x = aux[i]; ao |= x; aa &= x; }
    }
    for (int o = 16; o; o >>= 1)
    {
        ko |= __shfl_xor_sync(0xffffffffu, ko, o); ka &= __shfl_xor_sync(0xffffffffu, ka, o);
        ao |= __shfl_xor_sync(0xffffffffu, ao, o); aa &= __shfl_xor_sync(0xffffffffu, aa, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicOr((unsigned long long *)&out4[0], ko); atomicAnd((unsigned long long *)&out4[1], ka);
                                   atomicOr((unsigned long long *)&out4[2], ao); atomicAnd((unsigned long long *)&out4[3], aa); }
}
