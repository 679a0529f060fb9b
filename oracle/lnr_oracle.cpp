// TEST INFRASTRUCTURE ONLY -- see oracle/lnr_oracle.h.
//
// Sequential restatement of the `linear filter` approximate-map hot path. Reference citations are
// file:line under /root/reference. Deliberately simple: std::vector containers, libstdc++ std::sort
// with the reference's comparators (so the permutation of comparator ties is the reference's,
// SURVEY.md section 7 hard part 2), no parallelism except the outer per-read loop of orc_map_batch.
#include "lnr_oracle.h"
#include <algorithm>
#include <array>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>
#include <omp.h>

namespace {

typedef uint64_t u64;
typedef int64_t i64;
typedef std::pair<u64, u64> UPair;
typedef std::array<int, 3> int96;

// ------------------------------------------------------------------------------------------------
// 64-bit cord / anchor / hit encodings                           (include/cords.h:23-39, src/cords.cpp)
// ------------------------------------------------------------------------------------------------
const u64 ANCHOR_ZERO = 1ULL << 20;                 // cords.cpp:8  const_anchor_zero
const u64 F_END = 1ULL << 60;                       // cords.cpp:23 flagEnd
const u64 F_STRAND = 1ULL << 61;                    // cords.cpp:27 flag_strand
const u64 F_RECD = 1ULL << 62;                      // cords.cpp:36
const u64 F_MAIN = 1ULL << 63;                      // cords.cpp:35
const u64 MASK_Y = 0xfffff;                         // cords.cpp:24
const u64 MASK_X40 = 0xffffffffffULL;               // cords.cpp:25 maskx
const u64 VALUE_MASK = (1ULL << 60) - 1;            // cords.cpp:26
const u64 VALUE_MASK_DSTR = VALUE_MASK | F_STRAND;  // cords.cpp:33
const u64 MAX_CORD_ID = (1ULL << 10) - 1, MAX_CORD_X = (1ULL << 30) - 1;   // cords.cpp:13-14

inline u64 cord_x(u64 v) { return (v >> 20) & ((1ULL << 30) - 1); }        // cords.cpp:159
inline u64 cord_y(u64 v) { return v & MASK_Y; }                             // cords.cpp:163
inline u64 cord_strand(u64 v) { return (v >> 61) & 1ULL; }                  // cords.cpp:164
inline u64 cord_id(u64 v) { return (v >> 50) & ((1ULL << 10) - 1); }        // cords.cpp:166
inline u64 cord_x40(u64 v) { return (v >> 20) & MASK_X40; }                 // Cord::getCordX cords.cpp:49
inline bool is_end(u64 v) { return v & F_END; }                             // cords.cpp:281
inline void set_end(u64 & v) { v |= F_END; }
inline void unset_end(u64 & v) { v &= ~F_END; }
inline u64 create_cord(u64 id, u64 x, u64 y, u64 strand)                    // cords.cpp:195, :59
{
    return (((id << 30) + x) << 20) + y + (strand << 61);
}
inline u64 shift_cord(u64 v, i64 x, i64 y)                                  // cords.cpp:133
{
    return x < 0 ? v - (u64(-x) << 20) + u64(y) : v + (u64(x) << 20) + u64(y);
}
inline u64 hit2cord_dstr(u64 hit)                                           // cords.cpp:81
{
    u64 c = (hit + ((hit & MASK_Y) << 20) - (ANCHOR_ZERO << 20)) & VALUE_MASK_DSTR;
    return c & ~(1ULL << 62);
}
inline u64 anchor_x(u64 a) { return cord_x(hit2cord_dstr(a)); }             // getAnchorX cords.cpp:461
inline int cords_consecutive(u64 c1, u64 c2, u64 gap)                       // cords.cpp:306
{
    u64 x1 = cord_x(c1), x2 = cord_x(c2), y1 = cord_y(c1), y2 = cord_y(c2);
    return !cord_strand(c1 ^ c2) && x1 <= x2 && y1 <= y2 && x2 - x1 < gap && y2 - y1 < gap;
}
inline UPair forward_y(UPair se, u64 read_len)                              // getUPForwardy cords.cpp:469
{
    if (cord_strand(se.first)) return UPair(read_len - cord_y(se.second) - 1, read_len - cord_y(se.first) - 1);
    return UPair(cord_y(se.first), cord_y(se.second));
}
inline bool range_overlap(u64 a1, u64 a2, u64 b1, u64 b2) { return std::max(a1, b1) < std::min(a2, b2); }
inline bool cordy_overlap(u64 c11, u64 c12, u64 c21, u64 c22, u64 L)        // cords.cpp:450
{
    return cord_strand(c11 ^ c21)
               ? range_overlap(cord_y(c11), cord_y(c12), L - 1 - cord_y(c21), L - 1 - cord_y(c22))
               : range_overlap(cord_y(c11), cord_y(c12), cord_y(c21), cord_y(c22));
}

// ------------------------------------------------------------------------------------------------
// Rolling double-strand minimizer hash                                      (src/shape_extend.cpp)
// ------------------------------------------------------------------------------------------------
struct Shape
{
    unsigned span, weight;
    u64 h, crh, X, Y, strand;
    int leftChar, x;
    explicit Shape(unsigned s) : span(s), weight(s - 8), h(0), crh(0), X(0), Y(0), strand(0), leftChar(0), x(0) {}
};
inline u64 mask_bits(unsigned b) { return (1ULL << b) - 1; }

// shape_extend.cpp:86 -- skips forward past N so that `span` consecutive non-N bases follow
u64 hash_init(Shape & me, const uint8_t * it)
{
    me.leftChar = 0; me.h = 0; me.crh = 0; me.x = -3;
    u64 k = 0, count = 0;
    while (count < me.span)
    {
        if (it[k + count] == 4) { k += count + 1; count = 0; }
        else count++;
    }
    unsigned bit = 2;
    for (unsigned i = 0; i < me.span - 1; ++i)
    {
        u64 val = it[k + i];
        me.x += (int(val) << 1) - 3;
        me.h = (me.h << 2) + val;
        me.crh += ((3ULL - val) << bit);
        bit += 2;
    }
    return k;
}
// shape_extend.cpp:173
void hash_nexth(Shape & me, const uint8_t * it)
{
    u64 mask = mask_bits((me.span << 1) - 2);
    int v2 = it[me.span - 1];
    me.h = ((me.h & mask) << 2) + v2;
    me.crh = ((me.crh >> 2) & mask) + ((3ULL - (u64)(i64)v2) << ((me.span << 1) - 2));
    me.x += (v2 - me.leftChar) << 1;
    me.leftChar = it[0];
}
// shape_extend.cpp:245 (hashNextXX) + :282 (hashNextXY2) = hashNextX :341.  `it` points at the window start.
void hash_nextx(Shape & me, const uint8_t * it)
{
    u64 v1, v2, t = 0;
    unsigned span = me.span << 1, weight = me.weight << 1;
    if (me.x > 0) { v2 = me.h; me.strand = 0; }
    else { v2 = me.crh; me.strand = 1; }
    me.X = mask_bits(me.span << 1);
    for (unsigned k = 64 - span; k <= 64 - weight; k += 2)
    {
        v1 = v2 << k >> (64 - weight);
        if (me.X > v1) { me.X = v1; t = k; }
    }
    me.Y = 0;
    if (me.x > 0)
    {
        i64 d = (i64)(t >> 1) + me.span + me.weight - 32;
        for (i64 i = d; i < d + 4; i++)
        {
            i64 val = it[i];
            me.Y = val > 3 ? (me.Y << 2) : (me.Y << 2) + val;
        }
    }
    else
    {
        i64 d = -(i64)(t >> 1) - (i64)me.weight + 31;
        for (i64 i = d; i > d - 4; i--)
        {
            i64 val = 3 - (i64)it[i];
            me.Y = val < 0 ? (me.Y << 2) : (me.Y << 2) + val;
        }
    }
}
// shape_extend.cpp:132 hashNext (HIndex build: X plus residual-bit Y)
u64 hash_next_full(Shape & me, const uint8_t * it)
{
    u64 v1;
    unsigned t = 0, span = me.span << 1, weight = me.weight << 1;
    u64 v2 = it[me.span - 1];
    u64 mask = mask_bits(span - 2);
    me.h = ((me.h & mask) << 2) + v2;
    me.crh = ((me.crh >> 2) & mask) + ((3ULL - v2) << (span - 2));
    me.X = mask_bits(span);
    me.x += (int)((v2 - (u64)(i64)me.leftChar) << 1);
    me.leftChar = it[0];
    if (me.x > 0) { v2 = me.h; me.strand = 0; }
    else { v2 = me.crh; me.strand = 1; }
    for (unsigned k = 64 - span; k <= 64 - weight; k += 2)
    {
        v1 = v2 << k >> (64 - weight);
        if (me.X > v1) { me.X = v1; t = k; }
    }
    me.Y = (v2 >> (64 - t) << (64 - t - weight)) + (v2 & ((1ULL << (64 - t - weight)) - 1)) +
           ((u64)t << (span - weight - 1));
    return me.X;
}

// ------------------------------------------------------------------------------------------------
// Context
// ------------------------------------------------------------------------------------------------
struct Seq
{
    std::vector<uint8_t> s;   // dna5 bytes + 64 zero bytes of slack (out-of-range flank = 0)
    u64 len;
};
Seq make_seq(const uint8_t * p, u64 n)
{
    Seq q;
    q.len = n;
    q.s.assign(n + 64, 0);
    if (n) std::memcpy(q.s.data(), p, n);
    return q;
}

} // namespace

struct orc_ctx
{
    std::vector<Seq> genomes;
    int index_type, feature_type, threads, preset;
    float stop_ratio;
    // DIndex (include/index_util.h:99-120)
    std::vector<int32_t> dir;
    std::vector<u64> hs;
    // HIndex (include/index_util.h:139-248)
    std::vector<u64> ysa;
    u64 empty_dir;
    std::vector<u64> hkv;   // sorted (val1,val2) directory entries
    u64 htable_len;
    // genome features
    std::vector<std::vector<int96> > f2;
    std::vector<std::vector<int16_t> > f2s;
    // outputs
    std::vector<u64> out;
    std::vector<int32_t> outf;
    std::vector<int64_t> recs;   // cords2BamLink records, 8 values each
    std::vector<u64> cigs;       // cigar elements (operation << 32) | count
};

namespace {

// ------------------------------------------------------------------------------------------------
// DIndex build                                                         (src/index_util.cpp:1628-1803)
// ------------------------------------------------------------------------------------------------
void build_dindex(orc_ctx & c)
{
    const int64_t min_step = 8, max_step = 10, omit_block = 400;   // index_util.cpp:2551-2553
    const unsigned span = 21;
    const unsigned T = c.threads;
    c.dir.assign((1u << 26) + 1, 0);                                // DIndex::fullSize :1498
    struct Rec { u64 X; u64 val; };
    std::vector<Rec> recs;
    for (size_t gi = 0; gi < c.genomes.size(); gi++)
    {
        const uint8_t * s = c.genomes[gi].s.data();
        u64 len = c.genomes[gi].len;
        std::vector<i64> tb;
        for (unsigned j = 0; j < T; j++) tb.push_back(len / T * j);  // :1656
        tb.push_back((i64)len - span);                                // :1658
        for (unsigned t = 0; t < T; t++)                              // the omp parallel region :1659
        {
            i64 t_str = tb[t] + span, t_end = tb[t + 1] - span;
            i64 last_j = t_str - 1, count = 0;
            u64 preVal = ~0ULL;
            Shape sh(span);
            hash_init(sh, s + t_str);                                 // :1668 (return value ignored)
            for (i64 j = t_str; j < t_end; j++)
            {
                hash_nexth(sh, s + j);
                if (++count > min_step)
                {
                    hash_nextx(sh, s + j);
                    if (preVal != sh.X || j - last_j > max_step)
                    {
                        c.dir[sh.X]++;                                // :1682 atomicInc
                        Rec r;
                        r.X = sh.X;
                        r.val = create_cord(gi, j + ANCHOR_ZERO, sh.Y, sh.strand);   // :1763
                        recs.push_back(r);
                        preVal = sh.X;
                        last_j = j;
                    }
                    count = 0;
                }
            }
        }
    }
    // :1702-1721 omit large buckets, exclusive prefix sum
    i64 sum = 0;
    for (size_t i = 0; i < c.dir.size(); i++)
    {
        if (c.dir[i] > omit_block) c.dir[i] = 0;
        sum += c.dir[i];
        c.dir[i] = (int32_t)(sum - c.dir[i]);
    }
    c.hs.assign(sum, 0);
    // pass 2 (:1737-1781): records of non-empty buckets; slot order inside a bucket is irrelevant
    // because every bucket is sorted ascending afterwards (:1788-1796)
    std::vector<int32_t> fill(c.dir.begin(), c.dir.end());
    for (size_t i = 0; i < recs.size(); i++)
    {
        u64 X = recs[i].X;
        if (c.dir[X + 1] - c.dir[X]) c.hs[fill[X]++] = recs[i].val;
    }
    for (size_t i = 0; i + 1 < c.dir.size(); i++)
        if (c.dir[i + 1] - c.dir[i] > 1) std::sort(c.hs.begin() + c.dir[i], c.hs.begin() + c.dir[i + 1]);
}

// ------------------------------------------------------------------------------------------------
// 2-mer / 48-base features                                              (src/pmpfinder.cpp:484-652)
// ------------------------------------------------------------------------------------------------
const short infiN = 31;
const short units[25] = {   // pmpfinder.cpp:541-548
    0, 6, 12, 18, infiN,
    24, (1 << 8) + 0, (1 << 8) + 6, (1 << 8) + 12, infiN,
    (1 << 8) + 18, (1 << 8) + 24, (2 << 8) + 0, (2 << 8) + 6, infiN,
    (2 << 8) + 12, (2 << 8) + 18, (2 << 8) + 24, infiN, infiN,
    infiN, infiN, infiN, infiN, infiN};
inline void add2mer(int96 & v, const uint8_t * it)   // pmpfinder.cpp:549
{
    unsigned o = it[0] * 5 + it[1];
    unsigned i = units[o] >> 8;
    unsigned add = (1u << (units[o] & 255)) & ((1u << 31) - 1);
    v[i] += (int)add;
}
inline void inc96(int96 & a, const int96 & b) { a[0] += b[0]; a[1] += b[1]; a[2] += b[2]; }
inline void dec96(int96 & a, const int96 & b) { a[0] -= b[0]; a[1] -= b[1]; a[2] -= b[2]; }

// serial builder, used for reads (pmpfinder.cpp:556-588)
void features48_serial(const uint8_t * s, i64 n, std::vector<int96> & f)
{
    const int step = 16, w48 = 48;
    int96 zero = {0, 0, 0};
    std::vector<int96> buf(3, zero);
    f.assign((n - w48) / step + 1, zero);
    for (unsigned i = 0; i < 3; i++)
    {
        for (unsigned j = i << 4; j < (i << 4) + step; j++) add2mer(buf[i], s + j);
        inc96(f[0], buf[i]);
    }
    int next = 1, ii = 0;
    for (int i = step; i < n - w48 - 1; i += step)
    {
        f[next] = f[next - 1];
        dec96(f[next], buf[ii]);
        buf[ii] = zero;
        for (int j = i - step + w48; j < i + w48; j++) add2mer(buf[ii], s + j);
        inc96(f[next], buf[ii]);
        ii = (ii + 1) % 3;
        next++;
    }
    f.resize(next);
}
// parallel builder, used for genomes; `threads` only changes where streams restart, the values of
// the filled entries are chunk-independent (pmpfinder.cpp:589-652)
void features48_parallel(const uint8_t * s, i64 n, unsigned threads, std::vector<int96> & f)
{
    const int step = 16, w48 = 48;
    if (n < w48) { f.clear(); return; }
    int96 zero = {0, 0, 0};
    i64 range = (n - w48) / step + 1;
    if (range < (i64)threads) { features48_serial(s, n, f); return; }
    f.assign(((n - w48) >> 4) + 1, zero);
    for (unsigned t = 0; t < threads; t++)
    {
        i64 chunk = range / threads;
        i64 b = t * (chunk + 1);
        unsigned id1 = range - chunk * threads;
        if (t >= id1) b = id1 + chunk * t; else ++chunk;
        i64 e = b + chunk, next = b;
        b *= step; e *= step;
        std::vector<int96> buf(3, zero);
        f[next] = zero;
        for (unsigned i = 0; i < 3; i++)
        {
            unsigned tmp = b + (i << 4);
            for (unsigned j = tmp; j < tmp + step; j++) add2mer(buf[i], s + j);
            inc96(f[next], buf[i]);
        }
        int ii = 0;
        next++;
        for (int i = b + step; i < e; i += step)
        {
            f[next] = f[next - 1];
            dec96(f[next], buf[ii]);
            buf[ii] = zero;
            for (int j = i - step + w48; j < i + w48; j++) add2mer(buf[ii], s + j);
            inc96(f[next], buf[ii]);
            ii = (ii + 1) % 3;
            next++;
        }
    }
}

// pmpfinder.cpp:493-506 -- packed 6-bit fields, bias 31; borrows between fields are part of the spec
const int mxu31 = (31 << 24) + (31 << 18) + (31 << 12) + (31 << 6) + 31;
inline i64 script_dist(int s1, int s2)
{
    int d = s1 + mxu31 - s2;
    return std::abs((d >> 24 & 63) - 31) + std::abs((d >> 18 & 63) - 31) + std::abs((d >> 12 & 63) - 31) +
           std::abs((d >> 6 & 63) - 31) + std::abs((d & 63) - 31);
}
// _windowDist2_48 pmpfinder.cpp:523: scripts at feature offsets {0,3}
inline unsigned window_dist48(const int96 * a, const int96 * b)
{
    i64 sum = 0;
    for (unsigned i = 0; i < 6; i += 3)
        sum += script_dist(a[i][0], b[i][0]) + script_dist(a[i][1], b[i][1]) + script_dist(a[i][2], b[i][2]);
    return (unsigned)sum;
}

// base.cpp:335 _compltRvseStr  (complement table "tgcan")
void revcomp(const Seq & in, Seq & out)
{
    static const uint8_t cm[5] = {3, 2, 1, 0, 4};
    out.len = in.len;
    out.s.assign(in.len + 64, 0);
    for (u64 k = 0; k < in.len; k++) out.s[k] = cm[in.s[in.len - k - 1]];
}

// ------------------------------------------------------------------------------------------------
// Seeding against DIndex                                               (src/pmpfinder.cpp:1856-1911)
// ------------------------------------------------------------------------------------------------
void seed_dindex(const orc_ctx & c, const Seq & read, std::vector<u64> & set, u64 str, u64 end, int alpha)
{
    Shape sh(21);
    u64 L = read.len, xpre = 0;
    int dt = 0;
    const uint8_t * s = read.s.data();
    hash_init(sh, s);                                          // :1870 always at begin(read)
    for (u64 k = str + sh.span; k < end - sh.span; k++)        // :1874 (unsigned compare)
    {
        hash_nexth(sh, s + k);
        if (++dt == alpha)
        {
            dt = 0;
            hash_nextx(sh, s + k);
            if (sh.X ^ xpre)
            {
                i64 b = c.dir[sh.X], e = c.dir[sh.X + 1];      // queryHsStr/End index_util.cpp:2360
                for (i64 i = b; i < e; i++)
                {
                    u64 hsy = cord_y(c.hs[i]);
                    u64 val = hsy ^ sh.Y;
                    // `val >> ctz(val) < 4`; ctz(0) is UB in C, the oracle build accepts (App. C7)
                    if (val == 0 || (val >> __builtin_ctzl(val)) < 4)
                    {
                        u64 a;   // DIndex::val2Anchor index_util.cpp:1509
                        if (cord_strand(c.hs[i]) ^ sh.strand)
                        {
                            u64 cy = L - 1 - k;
                            a = (c.hs[i] - (cy << 20) + cy - hsy) | F_STRAND;
                        }
                        else
                            a = (c.hs[i] - (k << 20) + k - hsy) & ~F_STRAND;
                        set.push_back(a);
                    }
                }
                xpre = sh.X;
            }
        }
    }
}

} // namespace

#include "lnr_oracle_apx.inc"

// ------------------------------------------------------------------------------------------------
// C API
// ------------------------------------------------------------------------------------------------
extern "C" {

orc_ctx * orc_create(int n_contigs, const uint8_t * const * dna5, const uint64_t * lens,
                     int index_type, int feature_type, int threads, int preset, int build_index)
{
    orc_ctx * c = new orc_ctx();
    c->index_type = index_type;
    c->feature_type = feature_type;
    c->threads = threads;
    c->preset = preset;
    c->stop_ratio = preset == 0 ? 0.7f : 0.0f;   // mapper.cpp:174-197
    c->empty_dir = 0;
    c->htable_len = 0;
    for (int i = 0; i < n_contigs; i++) c->genomes.push_back(make_seq(dna5[i], lens[i]));
    c->f2.resize(n_contigs);
    c->f2s.resize(n_contigs);
    for (int i = 0; i < n_contigs; i++)
    {
        if (feature_type == 2) features48_parallel(c->genomes[i].s.data(), c->genomes[i].len, threads, c->f2[i]);
        else features32_parallel(c->genomes[i].s.data(), c->genomes[i].len, threads, c->f2s[i]);
    }
    if (build_index)
    {
        if (index_type == 1) build_dindex(*c);
        else build_hindex(*c);
    }
    return c;
}
void orc_destroy(orc_ctx * c) { delete c; }

int64_t orc_dindex_dir(orc_ctx * c, const int32_t ** p) { *p = c->dir.data(); return c->dir.size(); }
int64_t orc_dindex_hs(orc_ctx * c, const uint64_t ** p) { *p = c->hs.data(); return c->hs.size(); }
int64_t orc_hindex_ysa(orc_ctx * c, const uint64_t ** p, uint64_t * e)
{
    *p = c->ysa.data(); *e = c->empty_dir; return c->ysa.size();
}
int64_t orc_hindex_dir_kv(orc_ctx * c, const uint64_t ** p, uint64_t * tl)
{
    *p = c->hkv.data(); *tl = c->htable_len; return c->hkv.size() / 2;
}
int64_t orc_genome_features(orc_ctx * c, int contig, const int32_t ** p)
{
    c->outf.clear();
    if (c->feature_type == 2)
    {
        for (size_t i = 0; i < c->f2[contig].size(); i++)
            for (int k = 0; k < 3; k++) c->outf.push_back(c->f2[contig][i][k]);
        *p = c->outf.data();
        return c->f2[contig].size();
    }
    for (size_t i = 0; i < c->f2s[contig].size(); i++) c->outf.push_back(c->f2s[contig][i]);
    *p = c->outf.data();
    return c->f2s[contig].size();
}
int64_t orc_read_features(orc_ctx * c, const uint8_t * read, uint64_t len, int strand, const int32_t ** p)
{
    Seq r = make_seq(read, len), rc;
    revcomp(r, rc);
    const Seq & q = strand ? rc : r;
    c->outf.clear();
    int64_t n;
    if (c->feature_type == 2)
    {
        std::vector<int96> f;
        features48_serial(q.s.data(), q.len, f);
        for (size_t i = 0; i < f.size(); i++)
            for (int k = 0; k < 3; k++) c->outf.push_back(f[i][k]);
        n = f.size();
    }
    else
    {
        std::vector<int16_t> f;
        features32_serial(q.s.data(), q.len, f);
        for (size_t i = 0; i < f.size(); i++) c->outf.push_back(f[i]);
        n = f.size();
    }
    *p = c->outf.data();
    return n;
}

// ------------------------------------------------------------------------------------------------
// SAM* / BAM* record construction: cords2BamLink f_io.cpp:883, cord2cigar_ :758, ifCreateNew_ :673,
// createRectangleCigarPair :697, socreCigarPair :720, insertNewBamRecord align_util.cpp:301
// ------------------------------------------------------------------------------------------------
namespace {
struct Cig { char op; uint32_t count; };
struct BamRec { int32_t rid, begin_pos; uint32_t flag; int s1, s2, s3; std::vector<Cig> cigar; };
void append_shrink(std::vector<Cig> & c, const Cig & e)   // appendCigarShrink :658
{
    if (!c.empty() && c.back().op == e.op) c.back().count += e.count;
    else c.push_back(e);
}
void rect_pair(u64 cord1, u64 cord2, Cig & c1, Cig & c2, int f_m)   // createRectangleCigarPair :697 (uint64 differences, 32-bit counts)
{
    u64 dx = cord_x(cord2) - cord_x(cord1), dy = cord_y(cord2) - cord_y(cord1);
    c1.op = !f_m ? '=' : 'X';
    if (dx >= dy) { c2.op = 'D'; c1.count = (uint32_t)dy; c2.count = (uint32_t)(dx - dy); }
    else { c2.op = 'I'; c1.count = (uint32_t)dx; c2.count = (uint32_t)(dy - dx); }
}
void put_pair(std::vector<Cig> & cig, const Cig & c1, const Cig & c2)
{
    if (c1.count) append_shrink(cig, c1);
    if (c2.count) append_shrink(cig, c2);
}
u64 cord2cigar(u64 cigar_str, u64 c1s, u64 c1e, u64 c2s, BamRec & r, i64 thd_DI, i64 thd_X)   // cord2cigar_ :758
{
    Cig g1 = {0, 0}, g2 = {0, 0};
    u64 x0 = cord_x(cigar_str), y0 = cord_y(cigar_str), x11 = cord_x(c1s), y11 = cord_y(c1s);
    u64 x12 = cord_x(c1e), y12 = cord_y(c1e), x21 = cord_x(c2s), y21 = cord_y(c2s);
    if (x0 - y0 != x11 - y11) return ~0ULL;
    if (x12 < x21 && y12 < y21)
    {
        rect_pair(c1s, c1e, g1, g2, 0);
        put_pair(r.cigar, g1, g2);
        i64 DI = (i64)(x21 - x12 - y21 + y12);
        i64 X = (i64)std::min(x21 - x12, y21 - y12);
        if (std::abs(DI) > thd_DI && X > thd_X)
        {
            i64 split_n = std::min((i64)std::ceil(float(std::abs(DI)) / thd_DI), X);
            i64 split_DI = thd_DI, split_X = X / split_n;
            u64 str = c1e;
            rect_pair(c1e, c2s, g1, g2, 1);   // computed, not appended (:806)
            for (i64 i = 0; i < split_n - 1; i++)
            {
                u64 end = DI < 0 ? shift_cord(str, split_X, split_X + split_DI) : shift_cord(str, split_X + split_DI, split_X);
                rect_pair(str, end, g1, g2, 0);
                put_pair(r.cigar, g1, g2);
                str = end;
            }
            rect_pair(str, c2s, g1, g2, 1);
            put_pair(r.cigar, g1, g2);
        }
        else
        {
            rect_pair(c1e, c2s, g1, g2, 1);
            put_pair(r.cigar, g1, g2);
        }
    }
    else   // the three other quadrants (:777, :850, :857) all emit the rectangle cord1_str -> cord2_str
    {
        rect_pair(c1s, c2s, g1, g2, 0);
        put_pair(r.cigar, g1, g2);
    }
    // socreCigarPair :720 on the LAST pair built
    if ((g1.op == '=' || g1.op == 'X') && (g2.op == 'I' || g2.op == 'D'))
    {
        if (g1.op == '=') { r.s1 += g1.count; r.s3 += g1.count; }
        else r.s2 += g1.count;
        r.s2 += g2.count < 100 ? g2.count : 0;
        if (g2.op == 'I') r.s3 += g2.count;
    }
    return c2s;
}
void cords2bam(const u64 * cs, u64 n, u64 read_len, u64 window, u64 thd_large_X, i64 thd_DI, i64 thd_X, std::vector<BamRec> & out)
{
    const u64 d = (window << 20) | window;
    std::vector<int> rec_ptr, end_ptr;
    u64 cigar_str = 0;
    int f_new = 1; uint32_t flag = 0;
    for (u64 i = 1; i < n; i++)
    {
        if (f_new)
        {
            if (i != 1) { rec_ptr.push_back((int)out.size() - 1); end_ptr.push_back((int)i - 1); }
            f_new = 0;
            BamRec r;
            r.rid = (int32_t)cord_id(cs[i]); r.begin_pos = (int32_t)cord_x(cs[i]);
            r.flag = flag | (cord_strand(cs[i]) ? 16u : 0u);       // bam_flag_rvcmp align_util.cpp:5
            r.s1 = r.s2 = r.s3 = 0;
            int r_begin = (int)cord_y(cs[i]);
            if (r_begin != 0) r.cigar.push_back(Cig{'S', (uint32_t)r_begin});   // insertNewBamRecord :324 (f_soft = 1)
            out.push_back(r);
            cigar_str = cs[i];
            flag = 0;
        }
        u64 c1s = cs[i], c1e = cs[i] + d, c2s;
        bool last = i == n - 1;
        if (!last)
        {   // ifCreateNew_ :673
            u64 x11 = cord_x(cs[i]), y11 = cord_y(cs[i]), x12 = cord_x(c1e), y12 = cord_y(c1e), x21 = cord_x(cs[i + 1]), y21 = cord_y(cs[i + 1]);
            last = is_end(cs[i]) || x11 > x21 || y11 > y21 || ((i64)(x21 - x12) > (i64)thd_large_X && (i64)(y21 - y12) > (i64)thd_large_X) ||
                   cord_strand(cs[i] ^ cs[i + 1]);
        }
        if (last) { c2s = c1e; f_new = 1; flag = 2048; }         // bam_flag_suppl
        else c2s = cs[i + 1];
        cigar_str = cord2cigar(cigar_str, c1s, c1e, c2s, out.back(), thd_DI, thd_X);
        if (cigar_str == ~0ULL) break;
        if (i == n - 1) { rec_ptr.push_back((int)out.size() - 1); end_ptr.push_back((int)n - 1); }
    }
    for (size_t k = 0; k < end_ptr.size(); k++)
    {
        int clipped = (int)read_len - (int)cord_y(cs[end_ptr[k]] + d);
        if (clipped > 0) out[rec_ptr[k]].cigar.push_back(Cig{'S', (uint32_t)clipped});
    }
}
}  // namespace

int64_t orc_cords2bam(orc_ctx * c, uint64_t read_len, const uint64_t * cords, uint64_t n_cords, int window, uint64_t thd_large_X,
                      int64_t thd_DI, int64_t thd_X, const int64_t ** recs, const uint64_t ** cigars, uint64_t * n_cigars)
{
    std::vector<BamRec> out;
    cords2bam(cords, n_cords, read_len, (u64)window, thd_large_X, thd_DI, thd_X, out);
    c->recs.clear(); c->cigs.clear();
    for (size_t k = 0; k < out.size(); k++)
    {
        const BamRec & r = out[k];
        c->recs.push_back(r.rid); c->recs.push_back(r.begin_pos); c->recs.push_back(r.flag);
        c->recs.push_back(r.s1); c->recs.push_back(r.s2); c->recs.push_back(r.s3);
        c->recs.push_back((int64_t)c->cigs.size());
        for (size_t j = 0; j < r.cigar.size(); j++) c->cigs.push_back(((u64)(unsigned char)r.cigar[j].op << 32) | r.cigar[j].count);
        c->recs.push_back((int64_t)c->cigs.size());
    }
    *recs = c->recs.data(); *cigars = c->cigs.data(); *n_cigars = c->cigs.size();
    return (int64_t)out.size();
}

int64_t orc_read_stage(orc_ctx * c, const uint8_t * read, uint64_t len, int stage,
                       uint64_t str, uint64_t end, int toggle, const uint64_t ** p)
{
    ReadWork w(*c, read, len);
    c->out.clear();
    if (stage == 0)
    {
        std::vector<u64> cords;
        apx_map(*c, w, cords);
        c->out = cords;
    }
    else if (stage == 4)
    {
        std::vector<u64> cords;
        apx_map_(*c, w, cords, 0, create_cord(MAX_CORD_ID, MAX_CORD_X, len, 0));
        c->out = cords;
    }
    else
    {
        if (toggle) w.toggle(1);
        std::vector<u64> anchors(1, 0);   // Anchors::init(1) base.cpp:272
        seed(*c, w, anchors, str, create_cord(MAX_CORD_ID, MAX_CORD_X, end, 0));
        if (stage == 1) c->out = anchors;
        else if (stage == 2) { filter_anchors(anchors); c->out = anchors; }
        else if (stage == 3)
        {
            std::vector<u64> hits(1, F_END);
            std::vector<int> hits_score;
            anchor_hits_chains(w, anchors, hits, hits_score);
            c->out = hits;
        }
        else if (stage == 5)
        {
            std::vector<u64> hits(1, F_END);
            std::vector<int> hits_score(1, 0);
            filter_anchors(anchors);
            chain_anchors_hits(w, anchors, hits, hits_score);
            c->out = hits;
        }
    }
    *p = c->out.data();
    return (int64_t)c->out.size();
}

int orc_map_batch(orc_ctx * c, uint32_t n_reads, const uint8_t * bases, const uint64_t * read_off, int map_threads,
                  uint64_t * cords, uint64_t * cords_off, uint64_t cords_cap)
{
    std::vector<std::vector<u64> > res(n_reads);
#pragma omp parallel for schedule(dynamic, 4) num_threads(map_threads)
    for (uint32_t j = 0; j < n_reads; j++)
    {
        u64 len = read_off[j + 1] - read_off[j];
        if (len <= 200) continue;   // mapper.cpp:430,440
        ReadWork w(*c, bases + read_off[j], len);
        apx_map(*c, w, res[j]);
    }
    u64 tot = 0;
    cords_off[0] = 0;
    for (uint32_t j = 0; j < n_reads; j++)
    {
        if (tot + res[j].size() > cords_cap) return -1;
        if (!res[j].empty()) std::memcpy(cords + tot, res[j].data(), 8 * res[j].size());
        tot += res[j].size();
        cords_off[j + 1] = tot;
    }
    return 0;
}

// -c 0 (f_chain = 0): see apx_map_c0; gdl_state 0 = PMPParms as constructed, 1 = as left by an earlier toggle(0)
int orc_map_batch_c0(orc_ctx * c, uint32_t n_reads, const uint8_t * bases, const uint64_t * read_off, int map_threads, int gdl_state,
                     uint64_t * cords, uint64_t * cords_off, uint64_t cords_cap)
{
    std::vector<std::vector<u64> > res(n_reads);
#pragma omp parallel for schedule(dynamic, 4) num_threads(map_threads)
    for (uint32_t j = 0; j < n_reads; j++)
    {
        u64 len = read_off[j + 1] - read_off[j];
        if (len <= 200) continue;   // mapper.cpp:430,440
        ReadWork w(*c, bases + read_off[j], len);
        apx_map_c0(*c, w, res[j], gdl_state);
    }
    u64 tot = 0;
    cords_off[0] = 0;
    for (uint32_t j = 0; j < n_reads; j++)
    {
        if (tot + res[j].size() > cords_cap) return -1;
        if (!res[j].empty()) std::memcpy(cords + tot, res[j].data(), 8 * res[j].size());
        tot += res[j].size();
        cords_off[j + 1] = tot;
    }
    return 0;
}

} // extern "C"
