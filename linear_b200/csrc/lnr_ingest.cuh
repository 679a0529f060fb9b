// Read ingest on the device: FASTA / FASTQ text -> Dna5 ordinals back to back, read offsets and the
// span of every record id inside the text. Replaces, for the mapper's batches, seqan's readRecords + Dna5 conversion
// (loadRecords base.cpp:154, readRecords4FinPool2_ parallel_io.cpp:466; Dna5 table alphabet_residue_tabs.h:113-140:
// A/a C/c G/g T/t/U/u -> 0..3, every other byte -> 4). Same record model as the host reader of the CLI mirror
// (csrc/host/lnr_filter_main.cpp, which the APF parity tests pin against the reference binary):
//   FASTA: a line starting with '>' opens a record, its id is the rest of the line; every other line adds its bytes
//          except '\r' and ' ' to the record's sequence (line wrapping, CRLF and lower case are fine).
//   FASTQ: records of exactly four lines (id, sequence, '+', quality).
// The text must start with '>' or '@'.
//
// Passes (all bandwidth-bound; every thread owns 16 consecutive bytes of the text, which is read four times):
//   k_ing_count_nl     newlines per thread                                          -> scan -> line index of every thread
//   k_ing_line_starts  position after every newline (the line table)
//   k_ing_chunk_count  sequence bytes kept and records opened per thread            -> scans -> output offsets, record ids
//   k_ing_chunk_write  ordinals at the thread's output offset; offset + id span of every record it opens
// (A first version ran one warp per LINE for the last two passes: 80-column FASTA left most lanes idle, 1.06 of 1.4 ms.)
#pragma once

struct lnr_reads
{
    lnr_ctx * ctx = nullptr;
    u64 n_reads = 0, total_bases = 0, n_bytes = 0;
    int format = 0;              // 1 FASTA, 2 FASTQ
    u8 * d_bases = nullptr;      // total_bases (+256 zero bytes)
    u64 * d_off = nullptr;       // n_reads + 1
    u64 * d_id_off = nullptr;    // n_reads: offset of the id inside the text
    u32 * d_id_len = nullptr;    // n_reads
    std::vector<u64> h_off;      // host copy of d_off (batch layout of lnr_apxmap_reads)
    u8 * d_block = nullptr;      // the one allocation behind the four device arrays
    size_t block_bytes = 0;
};

__device__ __forceinline__ u32 ing_ord5(u32 c)
{
    c &= 0xdfu;   // upper case (only letters matter: every non-letter maps to N anyway)
    return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : (c == 'T' || c == 'U') ? 3u : 4u;
}

__global__ void __launch_bounds__(256) k_ing_count_nl(const u8 * __restrict__ text, u64 n, u32 * __restrict__ cnt, u64 n_threads)
{
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_threads) return;
    u64 b = t * 16, e = min(b + 16, n);
    u32 c = 0;
    if (e - b == 16 && (((uintptr_t)(text + b)) & 15) == 0)
    {
        uint4 v = __ldg((const uint4 *)(text + b));
        u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            u32 x = w[i] ^ 0x0a0a0a0au;   // zero bytes where '\n'
            u32 z = ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
            c += __popc(z);
        }
    }
    else
        for (u64 i = b; i < e; i++) c += text[i] == '\n';
    cnt[t] = c;
}
__global__ void __launch_bounds__(256) k_ing_line_starts(const u8 * __restrict__ text, u64 n, const u64 * __restrict__ off, u64 * __restrict__ ls,
                                                         u64 n_threads, u64 n_lines)
{
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) { ls[0] = 0; ls[n_lines] = n + 1; }
    if (t >= n_threads) return;
    u64 b = t * 16, e = min(b + 16, n);
    u64 k = 1 + off[t];
    for (u64 i = b; i < e; i++)
        if (text[i] == '\n') ls[k++] = i + 1;
}
// what a thread needs to know about the line its current byte is in
struct IngLine { bool hdr, seq; };
__device__ __forceinline__ IngLine ing_line(const u8 * __restrict__ text, const u64 * __restrict__ ls, u64 l, int format)
{
    IngLine r;
    const u64 b = ls[l], e = ls[l + 1] - 1;   // [b, e): the line without its newline
    if (format == 1) { r.hdr = e > b && text[b] == '>'; r.seq = !r.hdr; }
    else { r.hdr = (l & 3) == 0 && e > b; r.seq = (l & 3) == 1; }
    return r;
}
__device__ __forceinline__ bool ing_keep(u8 ch, int format) { return ch != '\r' && (format == 2 || ch != ' '); }   // the FASTQ reader drops only '\r'

template <bool WRITE>
__global__ void __launch_bounds__(256) k_ing_chunk(const u8 * __restrict__ text, u64 n, const u64 * __restrict__ ls, const u64 * __restrict__ line_of,
                                                   int format, u64 n_threads, u32 * __restrict__ kept, u32 * __restrict__ hdrs,
                                                   const u64 * __restrict__ kept_off, const u64 * __restrict__ rec_off, int cut_id_at_space,
                                                   u8 * __restrict__ bases, u64 * __restrict__ read_off, u64 * __restrict__ id_off, u32 * __restrict__ id_len)
{
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_threads) return;
    const u64 b = t * 16, e = min(b + 16, n);
    u8 ch[16];
    if (e - b == 16 && (((uintptr_t)(text + b)) & 15) == 0)
    {
        uint4 v = __ldg((const uint4 *)(text + b));
        u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 16; i++) ch[i] = (u8)(w[i >> 2] >> (8 * (i & 3)));
    }
    else
    {
#pragma unroll
        for (int i = 0; i < 16; i++) ch[i] = b + i < e ? text[b + i] : (u8)'\n';
    }
    u64 l = line_of[t];                  // line of the thread's first byte = newlines before it
    IngLine st = ing_line(text, ls, l, format);
    u32 c = 0, h = 0;
    u64 pos = WRITE ? kept_off[t] : 0, rec = WRITE ? rec_off[t] : 0;
#pragma unroll
    for (int i = 0; i < 16; i++)
    {
        const u64 p = b + i;
        if (p >= e) break;
        if (st.hdr && p == ls[l])
        {
            // this thread owns the first byte of a header line: it opens record `rec`
            if (WRITE)
            {
                u64 ib = p + 1, ie = ls[l + 1] - 1;
                if (ie > ib && text[ie - 1] == '\r') ie--;
                if (cut_id_at_space)
                    for (u64 k = ib; k < ie; k++) if (text[k] == ' ') { ie = k; break; }
                read_off[rec] = pos; id_off[rec] = ib; id_len[rec] = (u32)(ie - ib);
                rec++;
            }
            h++;
        }
        if (ch[i] == '\n') { l++; st = ing_line(text, ls, l, format); }
        else if (st.seq && ing_keep(ch[i], format))
        {
            if (WRITE) bases[pos++] = (u8)ing_ord5(ch[i]);
            c++;
        }
    }
    if (!WRITE) { kept[t] = c; hdrs[t] = h; }
}

static int reads_parse_device(lnr_ctx * ctx, const u8 * d_text, u64 n, int first_byte, int cut_id_at_space, lnr_reads ** out)
{
    const int format = first_byte == '>' ? 1 : first_byte == '@' ? 2 : 0;
    if (!format) return fail(ctx, LNR_E_ARG, "read text must start with '>' (FASTA) or '@' (FASTQ)");
    lnr_reads * R = new lnr_reads();
    R->ctx = ctx; R->n_bytes = n; R->format = format;
    *out = nullptr;
    // temporaries live in the context's reusable buffers (a cudaMalloc costs more than parsing tens of MB)
    u32 * d_cnt = nullptr; u64 * d_coff = nullptr, * d_ls = nullptr, * d_koff = nullptr, * d_rb = nullptr, * d_tot = nullptr;
    u32 * d_kept = nullptr, * d_hdr = nullptr;
    auto cleanup = [&]() {};
#define CKR(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); lnr_reads_destroy(R); return fail(ctx, LNR_E_CUDA, cudaGetErrorString(e_)); } } while (0)
    const u64 nt = (n + 15) / 16;
    CKR(ctx->ing[0].reserve((nt + STILE + 1) * sizeof(u32)));
    CKR(ctx->ing[1].reserve((nt + STILE + 1) * sizeof(u64)));
    CKR(ctx->ing[2].reserve(4 * sizeof(u64)));
    d_cnt = ctx->ing[0].as<u32>(); d_coff = ctx->ing[1].as<u64>(); d_tot = ctx->ing[2].as<u64>();
    {
        LaunchScope ls(ctx, "k_ing_count_nl");
        k_ing_count_nl<<<(u32)((nt + 255) / 256), 256, 0, ctx->stream>>>(d_text, n, d_cnt, nt);
    }
    int rc = device_scan<u64>(ctx, d_cnt, nt, 0, d_coff, d_tot, "k_ing_scan_nl");
    if (rc) { cleanup(); lnr_reads_destroy(R); return rc; }
    u64 n_nl = 0;
    CKR(cudaMemcpyAsync(&n_nl, d_tot, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CKR(cudaStreamSynchronize(ctx->stream));
    const u64 n_lines = n_nl + 1;
    CKR(ctx->ing[3].reserve((n_lines + 1) * sizeof(u64)));
    d_ls = ctx->ing[3].as<u64>();
    {
        LaunchScope ls(ctx, "k_ing_line_starts");
        k_ing_line_starts<<<(u32)((nt + 255) / 256), 256, 0, ctx->stream>>>(d_text, n, d_coff, d_ls, nt, n_lines);
    }
    CKR(ctx->ing[4].reserve((nt + STILE + 1) * sizeof(u32)));
    CKR(ctx->ing[5].reserve((nt + STILE + 1) * sizeof(u32)));
    CKR(ctx->ing[6].reserve((nt + STILE + 1) * sizeof(u64)));
    CKR(ctx->ing[7].reserve((nt + STILE + 1) * sizeof(u64)));
    d_kept = ctx->ing[4].as<u32>(); d_hdr = ctx->ing[5].as<u32>(); d_koff = ctx->ing[6].as<u64>(); d_rb = ctx->ing[7].as<u64>();
    const u32 chunk_ctas = (u32)((nt + 255) / 256);
    {
        LaunchScope ls(ctx, "k_ing_chunk_count");
        k_ing_chunk<false><<<chunk_ctas, 256, 0, ctx->stream>>>(d_text, n, d_ls, d_coff, format, nt, d_kept, d_hdr, nullptr, nullptr, 0, nullptr, nullptr,
                                                                nullptr, nullptr);
    }
    rc = device_scan<u64>(ctx, d_kept, nt, 0, d_koff, d_tot + 1, "k_ing_scan_kept");
    if (!rc) rc = device_scan<u64>(ctx, d_hdr, nt, 0, d_rb, d_tot + 2, "k_ing_scan_hdr");
    if (rc) { cleanup(); lnr_reads_destroy(R); return rc; }
    u64 tot[2] = {0, 0};
    CKR(cudaMemcpyAsync(tot, d_tot + 1, 2 * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CKR(cudaStreamSynchronize(ctx->stream));
    R->total_bases = tot[0]; R->n_reads = tot[1];
    {
        // one allocation: [read_off | id_off | id_len | bases + 256 zero bytes]
        const size_t nr = (size_t)R->n_reads + 1;
        const size_t o_idoff = nr * sizeof(u64), o_idlen = o_idoff + nr * sizeof(u64);
        const size_t o_bases = (o_idlen + nr * sizeof(u32) + 255) & ~(size_t)255;
        // (cudaMalloc + cudaFree of a block this size cost several times the parse itself: the context keeps the last
        // block a destroyed lnr_reads gave back and hands it out again when it is large enough)
        const size_t need = o_bases + R->total_bases + 256;
        u8 * blk = nullptr;
        if (ctx->reads_cache && ctx->reads_cache_bytes >= need)
        {
            blk = (u8 *)ctx->reads_cache; R->block_bytes = ctx->reads_cache_bytes;
            ctx->reads_cache = nullptr; ctx->reads_cache_bytes = 0;
        }
        else
        {
            CKR(cudaMalloc(&blk, need));
            R->block_bytes = need;
        }
        R->d_block = blk;
        R->d_off = (u64 *)blk; R->d_id_off = (u64 *)(blk + o_idoff); R->d_id_len = (u32 *)(blk + o_idlen); R->d_bases = blk + o_bases;
    }
    CKR(cudaMemsetAsync(R->d_bases + R->total_bases, 0, 256, ctx->stream));
    {
        LaunchScope ls(ctx, "k_ing_chunk_write");
        k_ing_chunk<true><<<chunk_ctas, 256, 0, ctx->stream>>>(d_text, n, d_ls, d_coff, format, nt, nullptr, nullptr, d_koff, d_rb, cut_id_at_space,
                                                               R->d_bases, R->d_off, R->d_id_off, R->d_id_len);
    }
    CKR(cudaMemcpyAsync(R->d_off + R->n_reads, &R->total_bases, sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    CKR(cudaGetLastError());
    R->h_off.resize(R->n_reads + 1);
    CKR(cudaMemcpyAsync(R->h_off.data(), R->d_off, (R->n_reads + 1) * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CKR(cudaStreamSynchronize(ctx->stream));
    cleanup();
#undef CKR
    *out = R;
    return LNR_OK;
}
