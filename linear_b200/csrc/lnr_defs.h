// Shared host/device definitions: the reference's 64-bit cord / anchor / hit encodings
// (include/cords.h:23-39, src/cords.cpp:21-37 of the reference) and small helpers.
// Compiles as CUDA device code and, for the CPU-side logic tests, as plain C++ (tests/host_emu).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define LNR_HD __host__ __device__ __forceinline__
// long, rarely executed bodies (introsort fallbacks, tracebacks): one out-of-line copy per instantiation instead of one
// per call site -- the warp-per-read kernels are bound by instruction fetch before anything else
#define LNR_HD_COLD __host__ __device__ __noinline__
#define LNR_DEV __device__
#else
#define LNR_HD inline
#define LNR_HD_COLD inline
#define LNR_DEV
#endif

namespace lnr {

typedef uint64_t u64;
typedef int64_t i64;
typedef uint32_t u32;
typedef int32_t i32;
typedef uint8_t u8;
typedef uint16_t u16;
typedef int16_t i16;

// ---- encodings -------------------------------------------------------------------------------------
static const u64 kAnchorZero = 1ULL << 20;            // const_anchor_zero (cords.cpp:8)
static const u64 kFlagEnd = 1ULL << 60;               // block end
static const u64 kFlagStrand = 1ULL << 61;
static const u64 kFlagRecd = 1ULL << 62;
static const u64 kFlagMain = 1ULL << 63;
static const u64 kMaskY = 0xfffffULL;
static const u64 kMaskX40 = 0xffffffffffULL;
static const u64 kValueMaskDstr = ((1ULL << 60) - 1) | (1ULL << 61);
static const u64 kMaxCordId = (1ULL << 10) - 1;
static const u64 kMaxCordX = (1ULL << 30) - 1;

LNR_HD u64 cord_x(u64 v) { return (v >> 20) & ((1ULL << 30) - 1); }       // get_cord_x cords.cpp:159
LNR_HD u64 cord_y(u64 v) { return v & kMaskY; }
LNR_HD u64 cord_strand(u64 v) { return (v >> 61) & 1ULL; }
LNR_HD u64 cord_id(u64 v) { return (v >> 50) & 1023ULL; }
LNR_HD u64 cord_x40(u64 v) { return (v >> 20) & kMaskX40; }               // Cord::getCordX (id|x)
LNR_HD bool is_end(u64 v) { return (v & kFlagEnd) != 0; }
LNR_HD u64 create_cord(u64 id, u64 x, u64 y, u64 strand) { return (((id << 30) + x) << 20) + y + (strand << 61); }
LNR_HD u64 shift_cord(u64 v, i64 x, i64 y)                                 // Cord::shift cords.cpp:133
{
    return x < 0 ? v - ((u64)(-x) << 20) + (u64)y : v + ((u64)x << 20) + (u64)y;
}
LNR_HD u64 hit2cord_dstr(u64 hit)                                          // cords.cpp:81
{
    return ((hit + ((hit & kMaskY) << 20) - (kAnchorZero << 20)) & kValueMaskDstr) & ~(1ULL << 62);
}
LNR_HD u64 anchor_x(u64 a) { return cord_x(hit2cord_dstr(a)); }            // getAnchorX cords.cpp:461
LNR_HD i64 iabs64(i64 v) { return v < 0 ? -v : v; }
LNR_HD i64 imax64(i64 a, i64 b) { return a > b ? a : b; }
LNR_HD i64 imin64(i64 a, i64 b) { return a < b ? a : b; }
LNR_HD u64 umin64(u64 a, u64 b) { return a < b ? a : b; }
LNR_HD u64 umax64(u64 a, u64 b) { return a > b ? a : b; }

LNR_HD int cords_consecutive(u64 c1, u64 c2, u64 gap)                      // isCordsConsecutive_ cords.cpp:306
{
    u64 x1 = cord_x(c1), x2 = cord_x(c2), y1 = cord_y(c1), y2 = cord_y(c2);
    return !cord_strand(c1 ^ c2) && x1 <= x2 && y1 <= y2 && x2 - x1 < gap && y2 - y1 < gap;
}

// ---- parameters fixed by the reference for this path ---------------------------------
static const int kSpanD = 21, kWeightD = 13;          // DIndex shape (index_util.cpp:2554, shape_extend.cpp:55)
static const int kDirBits = 26;                       // 2*weight
static const u32 kDirSize = (1u << 26) + 1;           // DIndex::fullSize index_util.cpp:1496
static const int kIdxMinStep = 8, kIdxMaxStep = 10, kIdxOmit = 400;   // index_util.cpp:2551-2553
static const int kMinReadLen = 200;                   // mapper.cpp:430
static const int kWin = 96, kSup = 6, kMed = 5, kInf = 3, kWinThr = 36, kWinReject = 50;   // ApxMapParm2_48
// ApxMapParm1_32 (-f 1, pmpfinder.cpp:199): cell 16 x 12 => window 192, sup 12, med ceil(.75 * 12), inf ceil(.5 * 12); same thresholds
static const int kWin32 = 192, kSup32 = 12, kMed32 = 9, kInf32 = 6;

// one int96 feature entry (std::array<int,3>, pmpfinder.h:71)
struct F96 { i32 v[3]; };

}  // namespace lnr
