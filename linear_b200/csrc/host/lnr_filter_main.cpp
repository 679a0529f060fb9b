// `linear_b200_filter filter reads.fa genome.fa [-t N] [-p N] [-i 1] [-f 2] [-c 1] [-ot 1] [-g 0]`
//
// Host-side mirror of the reference's CLI surface for the apx-map path (src/args_parser.cpp:118,
// src/linear.cpp:8-21 process1, src/mapper.cpp:883 map): load genome -> features + index on the GPU through the
// C ABI -> stream reads in blocks of 50 000 (mapper.cpp:892) -> lnr_apxmap_batch -> APF text identical to
// print_cords_apf (src/f_io.cpp:100-207). It exists to diff whole-program output against
// `linear filter reads.fa genome.fa -ot 1 -b 0 -g 0 -t T`; gap mapping and SAM/BAM stay the reference's host code.
// Output: <reads-file-stem>.apf in the working directory (mapper.cpp:904-906).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include "../../../include/lnr_b200.h"

struct Rec
{
    std::string id; std::vector<uint8_t> seq;
    uint64_t n = 0;   // sequence length when the bases live on the device only (GPU ingest)
    uint64_t length() const { return seq.empty() ? n : (uint64_t)seq.size(); }
};

static inline uint8_t ord5(char c)   // seqan Dna5: every non-ACGTU byte is N (alphabet_residue_tabs.h:113-140)
{
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2;
                 case 'T': case 't': case 'U': case 'u': return 3; default: return 4; }
}

// FASTA / FASTQ reader (--host-ingest); `n` records at most (0 = all). Returns false at end of file.
// Same record model as the device parser (lnr_ingest.cuh), which the APF tests pin against the reference binary: the
// first byte of the file decides the format. FASTA: only a line starting with '>' opens a record; every other line adds
// its bytes except '\r' and ' ' to the sequence. FASTQ: records of exactly four lines (a header line that is empty is
// skipped, with its three lines); '\r' is stripped from ids and sequences.
struct SeqReader
{
    std::ifstream in; std::string line; int format = 0; bool have = false;
    bool open(const std::string & p)
    {
        in.open(p);
        if (!in.good()) return false;
        const int c = in.peek();
        format = c == '>' ? 1 : (c == '@' ? 2 : 0);
        return true;
    }
    static void strip_cr(std::string & s) { if (!s.empty() && s.back() == '\r') s.pop_back(); }
    bool read(std::vector<Rec> & out, size_t n, bool cut_id_at_space)
    {
        out.clear();
        if (format == 1)
        {
            while (n == 0 || out.size() < n)
            {
                if (have) have = false;
                else if (!std::getline(in, line)) break;
                if (line.empty() || line[0] != '>') continue;       // bytes in front of the first record belong to nobody
                Rec r; r.id = line.substr(1);
                strip_cr(r.id);
                if (cut_id_at_space) r.id = r.id.substr(0, r.id.find(' '));   // loadRecords base.cpp:154
                while (std::getline(in, line))
                {
                    if (!line.empty() && line[0] == '>') { have = true; break; }
                    for (char c : line) if (c != '\r' && c != ' ') r.seq.push_back(ord5(c));
                }
                out.push_back(std::move(r));
            }
        }
        else if (format == 2)
        {
            std::string s, plus, q;
            while (n == 0 || out.size() < n)
            {
                if (!std::getline(in, line)) break;
                const bool more = (bool)std::getline(in, s);
                std::getline(in, plus); std::getline(in, q);
                if (line.empty()) continue;
                Rec r; r.id = line.substr(1);
                strip_cr(r.id);
                if (cut_id_at_space) r.id = r.id.substr(0, r.id.find(' '));
                if (more) for (char c : s) if (c != '\r') r.seq.push_back(ord5(c));
                out.push_back(std::move(r));
            }
        }
        return !out.empty();
    }
};

static const uint64_t kEnd = 1ULL << 60;
static inline uint64_t cx(uint64_t v) { return (v >> 20) & ((1ULL << 30) - 1); }
static inline uint64_t cy(uint64_t v) { return v & 0xfffff; }
static inline uint64_t cs(uint64_t v) { return (v >> 61) & 1; }
static inline uint64_t cid(uint64_t v) { return (v >> 50) & 1023; }

// print_cords_apf (f_io.cpp:100-207) for one block of reads
static void write_apf(std::ostream & of, const std::vector<Rec> & reads, const std::vector<Rec> & genome, const uint64_t * cords,
                      const uint64_t * coff, char & main_icon, uint64_t window)
{
    std::ostringstream st;
    int fflag = 0;
    for (size_t k = 0; k < reads.size(); k++)
    {
        const uint64_t * c = cords + coff[k];
        size_t n = (size_t)(coff[k + 1] - coff[k]);
        for (size_t j = 1; j < n; j++)
        {
            if (c[j - 1] & kEnd)
            {
                size_t m = j; int mcount = 0, blen = 0;
                while (m < n && !(c[m] & kEnd)) { if (cs(c[m])) mcount++; blen++; m++; }
                if (mcount > blen / 2) main_icon = '-';
                else if (mcount == blen / 2) main_icon = cs(c[j]) ? '-' : '+';
                else main_icon = '+';
                uint64_t rend = 0, gend = 0;
                for (size_t i = j;; i++)
                    if ((c[i] & kEnd) || i == n - 1) { rend = cy(c[i]) + window; gend = cx(c[i]) + window; break; }
                if (k > 0) st << "\n";
                st << "@ " << reads[k].id << " " << reads[k].length() << " " << cy(c[j]) << " "
                   << std::min<uint64_t>(rend, reads[k].length()) << " " << main_icon << " " << genome[cid(c[j])].id << " "
                   << genome[cid(c[j])].length() << " " << cx(c[j]) << " " << gend << "\n";
                fflag = 1;
            }
            char icon = cs(c[j]) ? '-' : '+';
            int64_t d1 = 0, d2 = 0;
            if (!fflag) { d1 = (int64_t)(cx(c[j]) - cx(c[j - 1])); d2 = (int64_t)(cy(c[j]) - cy(c[j - 1])); }
            st << "| " << cy(c[j]) << " " << cx(c[j]) << " " << d2 << " " << d1 << " " << icon << "\n";
            fflag = 0;
        }
    }
    of << st.str();
}

// ---- SAM* text (-ot 2): header (Mapper::setMapperBamHeaders mapper.cpp:288-321), then per read one line per record
// (printAlignSamRecord f_io.cpp:525 -> writeSam :313). The records come from lnr_cords_to_records; every record is its
// own line (cords2BamLink links nothing), a read with several records gets an SA:Z tag that lists the others
// (createSAZTagOneLine align_util.cpp:719, createSAZTagCigar :452 -- the trailing soft clip never reaches the
// simplified cigar there, :494, and zero-length elements are kept).
static void write_sam_header(std::ostream & of, const std::vector<Rec> & genome)
{
    for (auto & gnm : genome) of << "@SQ\tSN:" << gnm.id << "\tLN:" << gnm.length() << "\n";
    of << "@RG\tID:\tSM:\n" << "@PG\tID:M1-3\tPN:Linear\tCL:\n";
}
static void write_sam(std::ostream & of, const std::vector<Rec> & reads, const std::vector<Rec> & genome, const lnr_bam_rec * recs,
                      const uint64_t * rec_off, const uint64_t * cig, const uint64_t * cig_off)
{
    std::ostringstream st;
    for (size_t k = 0; k < reads.size(); k++)
    {
        const lnr_bam_rec * r = recs + rec_off[k];
        const size_t nr = (size_t)(rec_off[k + 1] - rec_off[k]);
        const uint64_t * c = cig + cig_off[k];
        // the chimeric entry of every record, used by the other records' tags. Its NM field is the record's edit count only
        // the first time the entry is produced, 0 afterwards: createSAZTagCigarOneChimeric (align_util.cpp:642) recomputes
        // nm_i only while the simplified cigar is not cached yet and overwrites record.nm_i with 0 on every later call
        std::vector<std::string> saz(nr), saz0(nr);
        if (nr > 1)
            for (size_t i = 0; i < nr; i++)
            {
                long cm = 0, ci = 0, nm = 0; uint32_t s0 = 0;
                for (uint32_t j = r[i].cigar_begin; j < r[i].cigar_end; j++)
                {
                    const char op = (char)(c[j] >> 32); const uint32_t cnt = (uint32_t)c[j];
                    if (j == r[i].cigar_begin && op == 'S') s0 = cnt;
                    else if (op == '=') cm += cnt;
                    else if (op == 'X') { cm += cnt; nm += cnt; }
                    else if (op == 'I') { ci -= cnt; nm += cnt; }
                    else if (op == 'D') { ci += cnt; nm += cnt; }
                }
                std::ostringstream e;
                e << genome[r[i].rid].id << "," << (r[i].begin_pos + 1) << "," << ((r[i].flag & 16) ? '-' : '+') << "," << s0 << "S" << (uint32_t)cm << "M"
                  << (uint32_t)(ci < 0 ? -ci : ci) << (ci < 0 ? 'I' : 'D') << "0S," << 255 << ",";
                saz[i] = e.str() + std::to_string(nm) + ";";
                saz0[i] = e.str() + "0;";
            }
        for (size_t i = 0; i < nr; i++)
        {
            st << reads[k].id << "\t" << r[i].flag << "\t" << genome[r[i].rid].id << "\t" << (r[i].begin_pos + 1) << "\t255\t";
            if (r[i].cigar_end == r[i].cigar_begin) st << "*";
            for (uint32_t j = r[i].cigar_begin; j < r[i].cigar_end; j++) st << (uint32_t)c[j] << (char)(c[j] >> 32);
            st << "\t*\t0\t0\t*\t*";
            if (nr > 1)
            {
                st << "\tSA:Z:";
                for (size_t j = 0; j < nr; j++)
                    if (j != i) st << (((j != 0 && i == 0) || (j == 0 && i == 1)) ? saz[j] : saz0[j]);
            }
            st << "\n";
        }
    }
    of << st.str();
}

int main(int argc, char ** argv)
{
    std::vector<std::string> pos;
    int threads = 16, preset = 1, index_t = 1, feature_t = 2, ot = 2, device = 0, host_ingest = 0;   // defaults: base.cpp:26-54
    int apx_chain = 1, gdl_state = 0;   // -c (apx_chain_flag, base.cpp:35); --gdl-state: lnr_params.gdl_state for -c 0
    for (int i = 1; i < argc; i++)
    {
        std::string a = argv[i];
        auto val = [&](int & dst) { if (i + 1 < argc) dst = atoi(argv[++i]); };
        if (a == "-t" || a == "--thread") val(threads);
        else if (a == "-p" || a == "--preset") val(preset);
        else if (a == "-i" || a == "--index_type") val(index_t);
        else if (a == "-f" || a == "--feature_type") val(feature_t);
        else if (a == "-ot" || a == "--output_type") val(ot);
        else if (a == "--device") val(device);
        else if (a == "--host-ingest") host_ingest = 1;
        else if (a == "-c" || a == "--apx_c_flag") val(apx_chain);
        else if (a == "--gdl-state") val(gdl_state);
        else if (a == "-g" || a == "-b" || a == "-o" || a == "-s" || a == "-a") { if (i + 1 < argc && argv[i + 1][0] != '-') i++; }
        else if (!a.empty() && a[0] == '-') { /* other reference options do not affect this path */ }
        else pos.push_back(a);
    }
    if (pos.size() < 3 || pos[0] != "filter")
    {
        fprintf(stderr, "usage: %s filter reads.fa genome.fa [-t N] [-p N] [-i 1|2] [-f 2|1] [-c 1|0] [-ot 1|2|3]\n", argv[0]);
        return 1;
    }
    const std::string rpath = pos[1], gpath = pos[2];
    lnr_ctx * ctx = nullptr; lnr_genome * g = nullptr; lnr_feats * f2 = nullptr; lnr_index * ix = nullptr;
    auto die = [&](const char * what, int rc) { fprintf(stderr, "lnr_b200: %s failed (%d): %s\n", what, rc, ctx ? lnr_last_error(ctx) : "no CUDA device"); return 2; };
    int rc;
    if ((rc = lnr_ctx_create(device, &ctx))) return die("lnr_ctx_create", rc);
    std::vector<Rec> genome;
    {
        // the genome takes the same road as the reads: bytes -> device parse -> lnr_genome_from_device (ids end at the
        // first blank, loadRecords base.cpp:154); the host reader is the fallback
        std::string gtext;
        std::ifstream gin(gpath, std::ios::binary);
        if (!gin.good()) { fprintf(stderr, "E[06]:Can't open file %s\n", gpath.c_str()); return 1; }
        if (!host_ingest) gtext.assign(std::istreambuf_iterator<char>(gin), std::istreambuf_iterator<char>());
        if (!host_ingest && !gtext.empty() && gtext[0] == '>')
        {
            lnr_reads * G = nullptr;
            if ((rc = lnr_reads_parse(ctx, gtext.data(), gtext.size(), 1, &G))) return die("lnr_reads_parse(genome)", rc);
            uint64_t nc = 0, tb = 0;
            lnr_reads_info(G, &nc, &tb);
            if (nc >= 1024) { fprintf(stderr, "too many contigs (linear.cpp:107)\n"); return 1; }
            std::vector<uint64_t> off(nc + 1), id_off(nc + 1), len(nc);
            std::vector<uint32_t> id_len(nc + 1);
            if ((rc = lnr_reads_download(G, nullptr, off.data(), id_off.data(), id_len.data()))) return die("lnr_reads_download(genome)", rc);
            genome.assign(nc, Rec());
            for (uint64_t i = 0; i < nc; i++)
            {
                genome[i].id.assign(gtext.data() + id_off[i], id_len[i]);
                genome[i].n = len[i] = off[i + 1] - off[i];
            }
            const uint8_t * dev_bases = nullptr;
            lnr_reads_device(G, &dev_bases, nullptr);
            if ((rc = lnr_genome_from_device(ctx, (uint32_t)nc, dev_bases, len.data(), &g))) return die("lnr_genome_from_device", rc);
            lnr_reads_destroy(G);
        }
        else
        {
            SeqReader gr;
            if (!gr.open(gpath)) { fprintf(stderr, "E[06]:Can't open file %s\n", gpath.c_str()); return 1; }
            gr.read(genome, 0, true);
            if (genome.size() >= 1024) { fprintf(stderr, "too many contigs (linear.cpp:107)\n"); return 1; }
            std::vector<const uint8_t *> ptr; std::vector<uint64_t> len;
            for (auto & r : genome) { ptr.push_back(r.seq.data()); len.push_back(r.seq.size()); }
            if ((rc = lnr_genome_upload(ctx, (uint32_t)genome.size(), ptr.data(), len.data(), &g))) return die("lnr_genome_upload", rc);
        }
    }
    if ((rc = lnr_features_build(ctx, g, feature_t, (unsigned)threads, &f2))) return die("lnr_features_build", rc);
    if ((rc = lnr_index_build(ctx, g, index_t, (unsigned)threads, &ix))) return die("lnr_index_build", rc);
    // output prefix = read file stem (mapper.cpp:904-906)
    std::string stem = rpath.substr(rpath.find_last_of('/') == std::string::npos ? 0 : rpath.find_last_of('/') + 1);
    stem = stem.substr(0, stem.find('.'));
    std::ofstream of, osam;
    if (ot & 1) of.open(stem + ".apf");
    if (ot & 2) { osam.open(stem + ".sam"); write_sam_header(osam, genome); }
    lnr_bam_parms bprm; memset(&bprm, 0, sizeof bprm);
    bprm.window = feature_t == 1 ? 192 : 96; bprm.thd_large_x = 8000;                           // mapper.cpp:466
    bprm.thd_di = preset == 1 ? 80 : (int64_t(1) << 60) - 1; bprm.thd_x = preset == 1 ? 200 : (int64_t(1) << 60) - 1;   // mapper.cpp:185, f_io.cpp:16
    auto emit_sam = [&](const std::vector<Rec> & rd, const uint64_t * cords, const uint64_t * coff) -> int {
        const uint32_t n = (uint32_t)rd.size();
        std::vector<uint64_t> len(n), roff(n + 1), goff(n + 1);
        for (uint32_t j = 0; j < n; j++) len[j] = rd[j].length();
        int rc2 = lnr_cords_to_records(ctx, n, cords, coff, len.data(), &bprm, nullptr, 0, roff.data(), nullptr, 0, goff.data());
        if (rc2 && rc2 != LNR_E_CAPACITY) return rc2;
        std::vector<lnr_bam_rec> recs(roff[n] + 1);
        std::vector<uint64_t> cig(goff[n] + 1);
        if (roff[n] && (rc2 = lnr_cords_to_records(ctx, n, cords, coff, len.data(), &bprm, recs.data(), roff[n], roff.data(), cig.data(), goff[n], goff.data()))) return rc2;
        write_sam(osam, rd, genome, recs.data(), roff.data(), cig.data(), goff.data());
        return 0;
    };
    lnr_params prm; memset(&prm, 0, sizeof prm); prm.preset = preset; prm.feature_type = feature_t;
    prm.no_chain = apx_chain == 0; prm.gdl_state = gdl_state;
    char main_icon = '+';
    uint64_t n_reads_total = 0, n_cords_total = 0;
    std::vector<Rec> reads;
    // read ingest on the device (lnr_reads_parse): the file's bytes are uploaded once, parsed there, and the batches are
    // mapped straight from the parsed buffers; --host-ingest (or a file that does not start with '>' / '@') takes the
    // host reader instead
    // The read file is streamed in chunks (default 2 GiB; a real read set is tens to hundreds of GB): a chunk's bytes are
    // uploaded once and parsed on the device, the complete blocks of `block_reads` reads (blockSize 50 000,
    // mapper.cpp:892) are mapped straight from the parsed buffers, and the next chunk starts at the first read that was
    // not mapped -- so the blocks, and with them the blank lines of the APF, are the same for every chunk size.
    uint64_t chunk_bytes = 2048ull << 20, block_reads = 50000;
    if (const char * e = getenv("LNR_CLI_CHUNK_MB")) { long v = atol(e); if (v >= 1) chunk_bytes = (uint64_t)v << 20; }
    if (const char * e = getenv("LNR_CLI_CHUNK_KB")) { long v = atol(e); if (v >= 1) chunk_bytes = (uint64_t)v << 10; }
    if (const char * e = getenv("LNR_CLI_BLOCK_READS")) { long v = atol(e); if (v >= 1) block_reads = (uint64_t)v; }
    std::string text;
    std::ifstream fin;
    uint64_t file_size = 0;
    if (!host_ingest)
    {
        fin.open(rpath, std::ios::binary);
        if (!fin.good()) { fprintf(stderr, "E[07]:Can't open read file %s\n", rpath.c_str()); return 1; }
        fin.seekg(0, std::ios::end);
        file_size = (uint64_t)fin.tellg();
        fin.seekg(0);
        const int c0 = fin.peek();
        if (file_size == 0 || (c0 != '>' && c0 != '@')) host_ingest = 1;
    }
    if (!host_ingest)
    {
        uint64_t pos = 0;
        while (pos < file_size)
        {
            const uint64_t want = std::min<uint64_t>(chunk_bytes, file_size - pos);
            text.resize(want);
            fin.clear();
            fin.seekg((std::streamoff)pos);
            fin.read(&text[0], (std::streamsize)want);
            const bool eof = pos + want >= file_size;
            lnr_reads * R = nullptr;
            if ((rc = lnr_reads_parse(ctx, text.data(), text.size(), 0, &R))) return die("lnr_reads_parse", rc);
            uint64_t n_all = 0, tb = 0;
            lnr_reads_info(R, &n_all, &tb);
            // the last record of a chunk may be cut off: it is never mapped from this chunk
            const uint64_t n_use = eof ? n_all : (n_all ? (n_all - 1) / block_reads * block_reads : 0);
            if (!eof && n_use == 0) { lnr_reads_destroy(R); chunk_bytes *= 2; continue; }   // not one whole block in the chunk
            std::vector<uint64_t> off(n_all + 1), id_off(n_all + 1);
            std::vector<uint32_t> id_len(n_all + 1);
            if ((rc = lnr_reads_download(R, nullptr, off.data(), id_off.data(), id_len.data()))) return die("lnr_reads_download", rc);
            for (uint64_t first = 0; first < n_use; first += block_reads)
            {
                const uint32_t n = (uint32_t)std::min<uint64_t>(block_reads, n_use - first);
                reads.assign(n, Rec());
                for (uint32_t j = 0; j < n; j++)
                {
                    reads[j].id.assign(text.data() + id_off[first + j], id_len[first + j]);
                    reads[j].n = off[first + j + 1] - off[first + j];
                }
                const uint64_t nb = off[first + n] - off[first];
                std::vector<uint64_t> cords(nb / 16 + 64 * (uint64_t)n + 1024), coff(n + 1);
                if ((rc = lnr_apxmap_reads(ctx, ix, f2, &prm, R, (uint32_t)first, n, cords.data(), coff.data(), cords.size(), nullptr)))
                    return die("lnr_apxmap_reads", rc);
                main_icon = '+';   // print_cords_apf re-initialises it per call (f_io.cpp:110)
                if (ot & 1) write_apf(of, reads, genome, cords.data(), coff.data(), main_icon, feature_t == 1 ? 192 : 96);
                if ((ot & 2) && (rc = emit_sam(reads, cords.data(), coff.data()))) return die("lnr_cords_to_records", rc);
                n_reads_total += n;
                n_cords_total += coff.back();
            }
            const uint64_t next = eof ? file_size : pos + id_off[n_use] - 1;   // the '>' / '@' in front of the first unmapped id
            lnr_reads_destroy(R);
            pos = next;
        }
    }
    else
    {
        SeqReader rr;
        if (!rr.open(rpath)) { fprintf(stderr, "E[07]:Can't open read file %s\n", rpath.c_str()); return 1; }
        while (rr.read(reads, 50000, false))   // blockSize, mapper.cpp:892
        {
            std::vector<uint64_t> off(reads.size() + 1, 0);
            for (size_t j = 0; j < reads.size(); j++) off[j + 1] = off[j] + reads[j].seq.size();
            std::vector<uint8_t> bases(off.back());
            for (size_t j = 0; j < reads.size(); j++) if (!reads[j].seq.empty()) memcpy(&bases[off[j]], reads[j].seq.data(), reads[j].seq.size());
            std::vector<uint64_t> cords(off.back() / 16 + 64 * reads.size() + 1024), coff(reads.size() + 1);
            if ((rc = lnr_apxmap_batch(ctx, ix, f2, &prm, (uint32_t)reads.size(), bases.data(), off.data(), cords.data(), coff.data(), cords.size(), nullptr)))
                return die("lnr_apxmap_batch", rc);
            main_icon = '+';   // print_cords_apf re-initialises it per call (f_io.cpp:110)
            if (ot & 1) write_apf(of, reads, genome, cords.data(), coff.data(), main_icon, feature_t == 1 ? 192 : 96);
            if ((ot & 2) && (rc = emit_sam(reads, cords.data(), coff.data()))) return die("lnr_cords_to_records", rc);
            n_reads_total += reads.size();
            n_cords_total += coff.back();
        }
    }
    fprintf(stderr, "lnr_b200 filter: %llu reads, %llu cords -> %s.apf\n", (unsigned long long)n_reads_total, (unsigned long long)n_cords_total, stem.c_str());
    lnr_index_destroy(ix); lnr_features_destroy(f2); lnr_genome_destroy(g); lnr_ctx_destroy(ctx);
    return 0;
}
