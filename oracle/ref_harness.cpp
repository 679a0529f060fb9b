// TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
//
// Function-level harness around the UNMODIFIED reference objects built from
// /root/reference by oracle/build_ref.sh. It is our own
// code: it only *calls* reference functions (all have external linkage) and exposes
// their inputs/outputs through a small C ABI so that tests/ and bench.py
// (cpu_baseline / --impl reference) can drive the real reference from Python (ctypes).
//
// Canonical-oracle conventions:
//   * private PMPParms per worker (semantics of `-b 1`),
//   * every read is copied into a buffer whose >= 16 bytes of slack past the end are zero
//     (the reference over-reads the forward Y flank by up to 3 bytes, shape_extend.cpp:292-297),
//   * `threads` is the reference's -t: it is a semantic parameter of the index
//     (index_util.cpp:1654-1670) and of the genome feature builder (pmpfinder.cpp:603-650).
//   * -f 1 (1-mer/32 features): the reference leaves the last 1-2 entries of a feature string unwritten
//     (createFeatures1_32 pmpfinder.cpp:354-373 serial, :393-423 parallel) and reads up to 10 entries past its end
//     (_windowDist checks the first index only :693-703; previousWindow has no upper bound :924). The canonical
//     definition adopted here -- the same way out-of-range flank bases read as 0 -- is: unwritten and out-of-range
//     feature entries are 0. The harness makes the reference follow it by giving every String<short> >= 64 entries of
//     capacity beyond its length BEFORE the reference resizes it (no reallocation), and zeroing everything from the
//     first unwritten entry on afterwards.
#include <omp.h>
#include <cstring>
#include <vector>
#include "base.h"
#include "cords.h"
#include "shape_extend.h"
#include "index_util.h"
#include "cluster_util.h"
#include "pmpfinder.h"
#include "align_util.h"

using namespace seqan;

// reference functions that have external linkage but no header declaration
unsigned getDIndexMatchAll(DIndex &, String<Dna5> &, String<uint64_t> &, uint64_t, uint64_t, PMPParms &);
unsigned getHIndexMatchAll(LIndex &, String<Dna5> &, String<uint64_t> &, uint64_t, uint64_t, PMPParms &);
uint64_t filterAnchors(Anchors &, uint64_t, uint64_t, uint64_t, unsigned, uint64_t, uint64_t, int);
int chainAnchorsHits(String<uint64_t> &, String<uint64_t> &, String<int> &, PMPParms &);
int getAnchorHitsChains(Anchors &, String<uint64_t> &, String<int> &, uint64_t, uint64_t, uint64_t, uint64_t,
                        unsigned, uint64_t, uint64_t, unsigned, GlobalParms &, PMPParms &);
uint64_t apxMap_(IndexDynamic &, String<Dna5> &, Anchors &, String<uint64_t> &, StringSet<FeaturesDynamic> &,
                 StringSet<FeaturesDynamic> &, String<uint64_t> &, String<CordInfo> &, uint64_t, uint64_t, int,
                 GlobalParms &, PMPParms &);
void _compltRvseStr(String<Dna5> & str, String<Dna5> & res);
int cords2BamLink(String<uint64_t> &, String<uint64_t> &, String<CordInfo> &, String<BamAlignmentRecordLink> &, String<Dna5> &, uint64_t,
                  int64_t, int64_t);   // f_io.cpp:883 (single read)

namespace {

struct RefCtx
{
    StringSet<String<Dna5> > genomes;
    StringSet<FeaturesDynamic> f2;
    IndexDynamic index;
    int feature_type;
    int threads;
    float stop_ratio;
    std::vector<uint64_t> out;   // last stage output
    std::vector<int32_t> outf;   // last feature output
    std::vector<uint64_t> kv;    // HIndex directory as sorted (key,val2) pairs
    std::vector<int64_t> recs;   // cords2BamLink records of the last call, 8 values each
    std::vector<uint64_t> cigs;  // their cigar elements, (operation << 32) | count
    RefCtx() : index(genomes) {}
};

// copy a dna5 byte string into a SeqAn string with zeroed slack behind the end
void fill_read(String<Dna5> & s, const uint8_t * p, uint64_t n)
{
    clear(s);
    reserve(s, n + 64, Exact());
    resize(s, n);
    std::memcpy((void *)&s[0], p, n);
    std::memset(((char *)&s[0]) + n, 0, 32);
}

// -f 1: first entry the reference's builders leave unwritten (serial: reads; parallel: genome contigs)
uint64_t f32_written_serial(uint64_t len)    // pmpfinder.cpp:354-373: f[0] + one entry per k = 16, 32, .. < len - 32
{
    if (len < 32) return 0;
    uint64_t lim = len - 32;
    return 1 + (lim > 16 ? (lim - 16 + 15) / 16 : 0);
}
uint64_t f32_written_parallel(uint64_t len)  // pmpfinder.cpp:393-423: entries [0, (len - 32 - 16) / 16)
{
    return len >= 48 ? (len - 48) / 16 : 0;
}
void f32_reserve(String<short> & f, uint64_t seq_len)
{
    uint64_t n = seq_len >= 32 ? ((seq_len - 32) >> 4) + 1 : 1;
    clear(f);
    reserve(f, n + 64, Exact());
}
void f32_zero_tail(String<short> & f, uint64_t written)
{
    uint64_t n = length(f);
    if (written > n) written = n;
    short * p = &f[0];
    std::memset(p + written, 0, (n + 64 - written) * sizeof(short));
}

struct Worker
{
    Anchors anchors;
    String<uint64_t> crhit;
    String<Dna5> read, comStr;
    String<UPair> apx_gaps;
    StringSet<FeaturesDynamic> f1;
    GlobalParms pm_g;
    PMPParms pm_pmp;
    Worker(RefCtx & c)
    {
        resize(f1, 2);
        f1[0].init(c.f2[0].fs_type);
        f1[1].init(c.f2[0].fs_type);
        pm_pmp.pm_cah.thd_stop_chain_len_ratio = c.stop_ratio;
    }
    void prepare(const uint8_t * p, uint64_t n)
    {
        fill_read(read, p, n);
        _compltRvseStr(read, comStr);
        reserve(comStr, n + 64, Exact());
        std::memset(((char *)&comStr[0]) + n, 0, 32);
        if (f1[0].isFs1_32()) { f32_reserve(f1[0].fs1_32, n); f32_reserve(f1[1].fs1_32, n); }
        createFeatures(begin(read), end(read), f1[0]);
        createFeatures(begin(comStr), end(comStr), f1[1]);
        if (f1[0].isFs1_32()) { f32_zero_tail(f1[0].fs1_32, f32_written_serial(n)); f32_zero_tail(f1[1].fs1_32, f32_written_serial(n)); }
    }
};

void copy_u64(std::vector<uint64_t> & dst, String<uint64_t> & src)
{
    dst.resize(length(src));
    if (length(src)) std::memcpy(dst.data(), &src[0], 8 * length(src));
}

} // namespace

extern "C" {

// index_type: 1 = DIndex (-i 1), 2 = HIndex (-i 2). feature_type: 2 = 2_48 (-f 2), 1 = 1_32 (-f 1).
// preset: -p (0 => stop ratio 0.7, 1/2 => 0; mapper.cpp:174-197).
void * ref_create(int n_contigs, const uint8_t * const * dna5, const uint64_t * lens,
                  int index_type, int feature_type, int threads, int preset, int build_index)
{
    RefCtx * c = new RefCtx();
    c->feature_type = feature_type;
    c->threads = threads;
    c->stop_ratio = preset == 0 ? 0.7f : 0.0f;
    resize(c->genomes, n_contigs);
    for (int i = 0; i < n_contigs; i++)
    {
        reserve(c->genomes[i], lens[i] + 64, Exact());
        resize(c->genomes[i], lens[i]);
        std::memcpy((void *)&c->genomes[i][0], dna5[i], lens[i]);
        std::memset(((char *)&c->genomes[i][0]) + lens[i], 0, 32);
    }
    omp_set_num_threads(threads);
    if (feature_type == 1)
    {
        resize(c->f2, n_contigs);
        for (int i = 0; i < n_contigs; i++) f32_reserve(c->f2[i].fs1_32, lens[i]);
    }
    createFeatures(c->genomes, c->f2, feature_type, (unsigned)threads);   // linear.cpp:14
    if (feature_type == 1)
        for (int i = 0; i < n_contigs; i++) f32_zero_tail(c->f2[i].fs1_32, f32_written_parallel(lens[i]));
    c->index.setIndexType(index_type);                                    // mapper.cpp:200
    if (build_index)
        createIndexDynamic(c->genomes, c->index, 0, n_contigs, threads, false);  // mapper.cpp:325
    return c;
}

void ref_destroy(void * h) { delete (RefCtx *)h; }

int64_t ref_dindex_dir(void * h, const int32_t ** p)
{
    RefCtx * c = (RefCtx *)h;
    String<int> & d = c->index.dindex.getDir();
    *p = length(d) ? (const int32_t *)&d[0] : 0;
    return length(d);
}
int64_t ref_dindex_hs(void * h, const uint64_t ** p)
{
    RefCtx * c = (RefCtx *)h;
    String<uint64_t> & d = c->index.dindex.getHs();
    *p = length(d) ? &d[0] : 0;
    return length(d);
}
int64_t ref_hindex_ysa(void * h, const uint64_t ** p, uint64_t * empty_dir)
{
    RefCtx * c = (RefCtx *)h;
    String<uint64_t> & d = c->index.hindex.ysa;
    *p = length(d) ? &d[0] : 0;
    *empty_dir = c->index.hindex.emptyDir;
    return length(d);
}
// HIndex directory as (val1,val2) pairs of all used slots, sorted (layout-independent view)
int64_t ref_hindex_dir_kv(void * h, const uint64_t ** p, uint64_t * table_len)
{
    RefCtx * c = (RefCtx *)h;
    XString & x = c->index.hindex.xstr;
    std::vector<std::pair<uint64_t, uint64_t> > v;
    for (uint64_t i = 0; i < length(x.xstring); i++)
        if (x.xstring[i].val1 != 0)   // val2 of empty slots is uninitialised (XString::_fullSize)
            v.push_back(std::make_pair((uint64_t)x.xstring[i].val1, (uint64_t)x.xstring[i].val2));
    std::sort(v.begin(), v.end());
    c->kv.clear();
    for (size_t i = 0; i < v.size(); i++) { c->kv.push_back(v[i].first); c->kv.push_back(v[i].second); }
    *p = c->kv.data();
    *table_len = length(x.xstring);
    return (int64_t)v.size();
}
// genome features of one contig: int96 (3 x int32 per entry) for -f 2, int16 for -f 1 (returned widened)
int64_t ref_genome_features(void * h, int contig, const int32_t ** p)
{
    RefCtx * c = (RefCtx *)h;
    FeaturesDynamic & f = c->f2[contig];
    c->outf.clear();
    if (f.isFs2_48())
    {
        for (unsigned i = 0; i < length(f.fs2_48); i++)
            for (int k = 0; k < 3; k++) c->outf.push_back(f.fs2_48[i][k]);
        *p = c->outf.data();
        return length(f.fs2_48);
    }
    for (unsigned i = 0; i < length(f.fs1_32); i++) c->outf.push_back(f.fs1_32[i]);
    *p = c->outf.data();
    return length(f.fs1_32);
}

// read features: strand 0 = read, 1 = reverse complement
int64_t ref_read_features(void * h, const uint8_t * read, uint64_t len, int strand, const int32_t ** p)
{
    RefCtx * c = (RefCtx *)h;
    Worker w(*c);
    w.prepare(read, len);
    FeaturesDynamic & f = w.f1[strand];
    c->outf.clear();
    int64_t n;
    if (f.isFs2_48())
    {
        n = length(f.fs2_48);
        for (unsigned i = 0; i < length(f.fs2_48); i++)
            for (int k = 0; k < 3; k++) c->outf.push_back(f.fs2_48[i][k]);
    }
    else
    {
        n = length(f.fs1_32);
        for (unsigned i = 0; i < length(f.fs1_32); i++) c->outf.push_back(f.fs1_32[i]);
    }
    *p = c->outf.data();
    return n;
}

// Stage checkpoints of one read:
//  stage 0: final cords_str of apxMap                       (pmpfinder.cpp:2709)
//  stage 1: raw anchors of the seeding call [str,end), `toggle`=0/1 (incl. anchors[0]=0 sentinel)
//  stage 2: anchors after filterAnchors (primary call)       (pmpfinder.cpp:2159)
//  stage 3: hits after getAnchorHitsChains (primary call)    (pmpfinder.cpp:2506)
//  stage 4: cords after the first apxMap_ (before clean_blocks_)
//  stage 5: hits after chainAnchorsHits only (before block stage)
int64_t ref_read_stage(void * h, const uint8_t * read, uint64_t len, int stage,
                       uint64_t str, uint64_t end, int toggle, const uint64_t ** p)
{
    RefCtx * c = (RefCtx *)h;
    Worker w(*c);
    w.prepare(read, len);
    c->out.clear();
    if (stage == 0)
    {
        String<uint64_t> cords_str, cords_end;
        String<CordInfo> cords_info;
        apxMap(c->index, w.read, w.anchors, w.crhit, w.f1, c->f2, w.apx_gaps, cords_str, cords_end, cords_info,
               1, w.pm_g, w.pm_pmp);
        copy_u64(c->out, cords_str);
    }
    else if (stage == 4)
    {
        String<uint64_t> cords_str;
        String<CordInfo> cords_info;
        uint64_t map_end = create_cord(MAX_CORD_ID, MAX_CORD_X, length(w.read), 0);
        apxMap_(c->index, w.read, w.anchors, w.crhit, w.f1, c->f2, cords_str, cords_info, 0, map_end, 2,
                w.pm_g, w.pm_pmp);
        copy_u64(c->out, cords_str);
    }
    else
    {
        w.anchors.init(1);
        if (toggle) w.pm_pmp.toggle(1);
        if (c->index.isHIndex())
        {
            uint64_t map_end = create_cord(MAX_CORD_ID, MAX_CORD_X, end, 0);
            getHIndexMatchAll(c->index.hindex, w.read, w.anchors.set, str, map_end, w.pm_pmp);
        }
        else
            getDIndexMatchAll(c->index.dindex, w.read, w.anchors.set, str, end, w.pm_pmp);
        if (stage == 1) copy_u64(c->out, w.anchors.set);
        else if (stage == 2)
        {
            filterAnchors(w.anchors, w.pm_g.shape_len, 1, 2, 2, 5, 2500, 2);
            copy_u64(c->out, w.anchors.set);
        }
        else if (stage == 3)
        {
            String<uint64_t> hits;
            String<int> hits_score;
            initHits(hits);
            getAnchorHitsChains(w.anchors, hits, hits_score, len, 1, 2, 600, 2, 5, 2500, 2, w.pm_g, w.pm_pmp);
            copy_u64(c->out, hits);
        }
        else if (stage == 5)
        {
            String<uint64_t> hits;
            String<int> hits_score;
            initHits(hits);
            filterAnchors(w.anchors, w.pm_g.shape_len, 1, 2, 2, 5, 2500, 2);
            initHitsScore(hits_score);
            chainAnchorsHits(w.anchors.set, hits, hits_score, w.pm_pmp);
            copy_u64(c->out, hits);
        }
    }
    *p = c->out.data();
    return (int64_t)c->out.size();
}

// SAM* / BAM* record construction: cords2BamLink (f_io.cpp:883) for one read. cords_end[i] =
// cords_str[i] + ((window << 20) | window) as apxMap leaves them. Returns the number of records; *recs holds 8 values per
// record (rID, beginPos, flag, score.s1, s2, s3, first cigar element, one past the last), *cigars the elements.
int64_t ref_cords2bam(void * h, uint64_t read_len, const uint64_t * cords, uint64_t n_cords, int window, uint64_t thd_large_X,
                      int64_t thd_DI, int64_t thd_X, const int64_t ** recs, const uint64_t ** cigars, uint64_t * n_cigars)
{
    RefCtx * c = (RefCtx *)h;
    String<uint64_t> cs, ce;
    String<CordInfo> ci;
    String<BamAlignmentRecordLink> recs_;
    String<Dna5> read;
    resize(read, read_len);                       // only its length is read (f_io.cpp:989)
    resize(cs, n_cords); resize(ce, n_cords);
    const uint64_t d = ((uint64_t)window << 20) | (uint64_t)window;
    for (uint64_t i = 0; i < n_cords; i++) { cs[i] = cords[i]; ce[i] = cords[i] + d; }
    cords2BamLink(cs, ce, ci, recs_, read, thd_large_X, thd_DI, thd_X);
    c->recs.clear(); c->cigs.clear();
    for (unsigned k = 0; k < length(recs_); k++)
    {
        BamAlignmentRecordLink & r = recs_[k];
        c->recs.push_back(r.rID); c->recs.push_back(r.beginPos); c->recs.push_back(r.flag);
        c->recs.push_back(r.score.s1); c->recs.push_back(r.score.s2); c->recs.push_back(r.score.s3);
        c->recs.push_back((int64_t)c->cigs.size());
        for (unsigned j = 0; j < length(r.cigar); j++)
            c->cigs.push_back(((uint64_t)(unsigned char)r.cigar[j].operation << 32) | (uint64_t)r.cigar[j].count);
        c->recs.push_back((int64_t)c->cigs.size());
    }
    *recs = c->recs.data(); *cigars = c->cigs.data(); *n_cigars = c->cigs.size();
    return (int64_t)length(recs_);
}

// Batch apx-map (the reference's own per-read body, mapper.cpp:438-447 without mapGaps), `map_threads`
// OpenMP workers with private scratch/params. cords_off must hold n_reads+1 entries; cords (capacity
// cords_cap) receives cords_str of every read back to back. Returns 0, or -1 if capacity is too small.
int ref_map_batch(void * h, uint32_t n_reads, const uint8_t * bases, const uint64_t * read_off, int map_threads,
                  uint64_t * cords, uint64_t * cords_off, uint64_t cords_cap)
{
    RefCtx * c = (RefCtx *)h;
    std::vector<std::vector<uint64_t> > res(n_reads);
    omp_set_num_threads(map_threads);
#pragma omp parallel
    {
        Worker w(*c);
#pragma omp for schedule(dynamic, 4)
        for (uint32_t j = 0; j < n_reads; j++)
        {
            uint64_t len = read_off[j + 1] - read_off[j];
            if (len <= 200) continue;   // mapper.cpp:430,440
            w.prepare(bases + read_off[j], len);
            String<uint64_t> cords_str, cords_end;
            String<CordInfo> cords_info;
            apxMap(c->index, w.read, w.anchors, w.crhit, w.f1, c->f2, w.apx_gaps, cords_str, cords_end,
                   cords_info, 1, w.pm_g, w.pm_pmp);
            copy_u64(res[j], cords_str);
        }
    }
    omp_set_num_threads(c->threads);
    uint64_t tot = 0;
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

// -c 0 (f_chain = 0, alg_type 1; apxMap pmpfinder.cpp:2773-2787). The reference keeps ONE PMPParms per thread and this
// mode leaves it changed: GetDHitListParms is constructed in its toggle(1) state (thd_list_n 20, thd_best_n 1, :2290-2297)
// and the first read that needs the second attempt ends with toggle(0) (10, 999) for every later read of that thread.
// Canonical convention: a fresh PMPParms per read; gdl_state 1 applies toggle(0) first (the "later reads" state).
int ref_map_batch_c0(void * h, uint32_t n_reads, const uint8_t * bases, const uint64_t * read_off, int map_threads, int gdl_state,
                     uint64_t * cords, uint64_t * cords_off, uint64_t cords_cap)
{
    RefCtx * c = (RefCtx *)h;
    std::vector<std::vector<uint64_t> > res(n_reads);
    omp_set_num_threads(map_threads);
#pragma omp parallel
    {
        Worker w(*c);
#pragma omp for schedule(dynamic, 4)
        for (uint32_t j = 0; j < n_reads; j++)
        {
            uint64_t len = read_off[j + 1] - read_off[j];
            if (len <= 200) continue;   // mapper.cpp:430,440
            w.prepare(bases + read_off[j], len);
            PMPParms fresh;
            fresh.pm_cah.thd_stop_chain_len_ratio = c->stop_ratio;
            if (gdl_state) fresh.toggle(0);
            String<uint64_t> cords_str, cords_end;
            String<CordInfo> cords_info;
            apxMap(c->index, w.read, w.anchors, w.crhit, w.f1, c->f2, w.apx_gaps, cords_str, cords_end,
                   cords_info, 0, w.pm_g, fresh);
            copy_u64(res[j], cords_str);
        }
    }
    omp_set_num_threads(c->threads);
    uint64_t tot = 0;
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
