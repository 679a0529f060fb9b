// SAM* / BAM* record construction from cords: cords2BamLink f_io.cpp:883, cord2cigar_ :758,
// ifCreateNew_ :673, createRectangleCigarPair :697, socreCigarPair :720, insertNewBamRecord align_util.cpp:301.
//
// One walk over a read's cords emits its records (contig, position, flag, score) and their cigar* elements. The walk is a
// template over a sink so that the same code counts (pass 1) and fills (pass 2) on the device -- one thread per read,
// exact-size output through a device scan, like the anchors -- and runs on the host in tests/host_emu.
// A record's trailing soft clip is appended when the record closes: the reference appends it in a second loop
// (:986-999), but nothing else is ever added to a closed record, so the cigar is the same.
#pragma once
#include "lnr_defs.h"

namespace lnr {

struct BamParms
{
    u32 window;        // cords_end[i] = cords_str[i] + ((window << 20) | window): 96 for -f 2, 192 for -f 1
    u64 thd_large_x;   // 8000 (mapper.cpp:466)
    i64 thd_di, thd_x; // FIOParms::thd_DI / thd_X: 2^60 - 1 by default, 80 / 200 with -p 0 (f_io.cpp:16, mapper.cpp:185)
};
struct BamRec { i32 rid, begin_pos; u32 flag; i32 s1, s2, s3; u32 cigar_begin, cigar_end; };   // cigar range relative to the read's first element

// pass 1: sizes only. Adjacent equal operations merge (appendCigarShrink :658), so only the last operation matters.
struct BamCountSink
{
    u32 n_rec = 0, n_cig = 0; int last_op = 0;
    LNR_HD void begin(i32, i32, u32) { n_rec++; last_op = 0; }
    LNR_HD void raw(int op, u32) { n_cig++; last_op = op; }
    LNR_HD void shrink(int op, u32) { if (op != last_op) { n_cig++; last_op = op; } }
    LNR_HD void close(i32, i32, i32) {}
};
// pass 2: writes records and elements; the element being merged into lives in registers until the next one arrives
struct BamFillSink
{
    BamRec * recs; u64 * cig; u32 n_rec = 0, n_cig = 0; int pend_op = 0; u32 pend_cnt = 0; BamRec cur;
    LNR_HD void flush() { if (pend_op) { cig[n_cig++] = ((u64)(u32)pend_op << 32) | pend_cnt; pend_op = 0; pend_cnt = 0; } }
    LNR_HD void begin(i32 rid, i32 pos, u32 flag) { cur.rid = rid; cur.begin_pos = pos; cur.flag = flag; cur.cigar_begin = n_cig; }
    LNR_HD void raw(int op, u32 c) { flush(); pend_op = op; pend_cnt = c; flush(); }
    LNR_HD void shrink(int op, u32 c) { if (op == pend_op) pend_cnt += c; else { flush(); pend_op = op; pend_cnt = c; } }
    LNR_HD void close(i32 s1, i32 s2, i32 s3) { flush(); cur.s1 = s1; cur.s2 = s2; cur.s3 = s3; cur.cigar_end = n_cig; recs[n_rec++] = cur; }
};

struct CigPair { int op1, op2; u32 c1, c2; };
LNR_HD CigPair rect_pair(u64 cord1, u64 cord2, int f_m)   // createRectangleCigarPair: uint64 differences, 32-bit counts
{
    const u64 dx = cord_x(cord2) - cord_x(cord1), dy = cord_y(cord2) - cord_y(cord1);
    CigPair p;
    p.op1 = !f_m ? '=' : 'X';
    if (dx >= dy) { p.op2 = 'D'; p.c1 = (u32)dy; p.c2 = (u32)(dx - dy); }
    else { p.op2 = 'I'; p.c1 = (u32)dx; p.c2 = (u32)(dy - dx); }
    return p;
}
template <class Sink> LNR_HD void put_pair(Sink & out, const CigPair & p)
{
    if (p.c1) out.shrink(p.op1, p.c1);
    if (p.c2) out.shrink(p.op2, p.c2);
}

template <class Sink>
LNR_HD void bam_walk(const u64 * cs, u32 n, u64 read_len, const BamParms & P, Sink & out)
{
    const u64 d = ((u64)P.window << 20) | (u64)P.window;
    bool f_new = true, open = false;
    u32 flag = 0;
    i32 s1 = 0, s2 = 0, s3 = 0;
    for (u32 i = 1; i < n; i++)
    {
        const u64 c1s = cs[i], c1e = c1s + d;
        if (f_new)
        {
            f_new = false;
            out.begin((i32)cord_id(c1s), (i32)cord_x(c1s), flag | (cord_strand(c1s) ? 16u : 0u));   // bam_flag_rvcmp
            open = true;
            s1 = s2 = s3 = 0;
            const i32 r_begin = (i32)cord_y(c1s);
            if (r_begin != 0) out.raw('S', (u32)r_begin);     // insertNewBamRecord: leading soft clip
            flag = 0;
        }
        bool last = i == n - 1;
        u64 c2s = 0;
        if (!last)
        {
            c2s = cs[i + 1];
            const u64 x11 = cord_x(c1s), y11 = cord_y(c1s), x12 = cord_x(c1e), y12 = cord_y(c1e), x21 = cord_x(c2s), y21 = cord_y(c2s);
            last = is_end(c1s) || x11 > x21 || y11 > y21 || ((i64)(x21 - x12) > (i64)P.thd_large_x && (i64)(y21 - y12) > (i64)P.thd_large_x) ||
                   cord_strand(c1s ^ c2s);                                                                // ifCreateNew_
        }
        if (last) { c2s = c1e; f_new = true; flag = 2048; }   // bam_flag_suppl for the record that follows
        // cord2cigar_ (cigar_str == cord1_str always holds here, so its error exit is unreachable)
        const u64 x12 = cord_x(c1e), y12 = cord_y(c1e), x21 = cord_x(c2s), y21 = cord_y(c2s);
        CigPair g;
        if (x12 < x21 && y12 < y21)
        {
            g = rect_pair(c1s, c1e, 0);
            put_pair(out, g);
            const i64 DI = (i64)(x21 - x12 - y21 + y12);
            const i64 X = (i64)umin64(x21 - x12, y21 - y12);
            const i64 aDI = DI < 0 ? -DI : DI;
            if (aDI > P.thd_di && X > P.thd_x)
            {
                // split_n = min(ceil(float(|DI|) / thd_DI), X): float division of a float by the int64 converted to float
                const float q = (float)aDI / (float)P.thd_di;
                i64 cq = (i64)q; if ((float)cq < q) cq++;
                const i64 split_n = cq < X ? cq : X;
                const i64 split_x = X / split_n;
                u64 str = c1e;
                for (i64 k = 0; k < split_n - 1; k++)
                {
                    const u64 end = DI < 0 ? shift_cord(str, split_x, split_x + P.thd_di) : shift_cord(str, split_x + P.thd_di, split_x);
                    g = rect_pair(str, end, 0);
                    put_pair(out, g);
                    str = end;
                }
                g = rect_pair(str, c2s, 1);
                put_pair(out, g);
            }
            else
            {
                g = rect_pair(c1e, c2s, 1);
                put_pair(out, g);
            }
        }
        else
        {
            g = rect_pair(c1s, c2s, 0);
            put_pair(out, g);
        }
        // socreCigarPair on the last pair built (op2 is always 'I' or 'D')
        if (g.op1 == '=') { s1 += (i32)g.c1; s3 += (i32)g.c1; }
        else s2 += (i32)g.c1;
        s2 += g.c2 < 100u ? (i32)g.c2 : 0;
        if (g.op2 == 'I') s3 += (i32)g.c2;
        if (last)
        {
            const i32 clipped = (i32)read_len - (i32)cord_y(c1e);
            if (clipped > 0) out.raw('S', (u32)clipped);
            out.close(s1, s2, s3);
            open = false;
        }
    }
    (void)open;
}

}  // namespace lnr
