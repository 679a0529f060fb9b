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
// Passes (all bandwidth-bound, the text is read four times):
//   k_ing_count_nl     thread = 16 bytes: newlines per thread                       -> scan -> line starts
//   k_ing_line_starts  thread = 16 bytes: position after every newline
//   k_ing_line_info    warp = one line : header flag, bytes kept                    -> scans -> output offsets, record ids
//   k_ing_write        warp = one line : ordinals (ballot compaction inside the line) / offset + id span of a record
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
// kept[l] = sequence bytes line l contributes; hdr[l] = 1 when it opens a record
__global__ void __launch_bounds__(256) k_ing_line_info(const u8 * __restrict__ text, const u64 * __restrict__ ls, u64 n_lines, int format,
                                                       u32 * __restrict__ kept, u32 * __restrict__ hdr)
{
    u64 l = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31;
    if (l >= n_lines) return;
    const u64 b = ls[l], e = ls[l + 1] - 1;   // [b, e): the line without its newline
    bool is_hdr, is_seq;
    if (format == 1) { is_hdr = e > b && text[b] == '>'; is_seq = !is_hdr; }
    else { is_hdr = (l & 3) == 0 && e > b; is_seq = (l & 3) == 1; }
    u32 c = 0;
    if (is_seq)
        for (u64 i = b + lane; i < e; i += 32) { u8 ch = text[i]; c += (ch != '\r' && (format == 2 || ch != ' ')) ? 1u : 0u; }
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) { kept[l] = c; hdr[l] = is_hdr ? 1u : 0u; }
}
__global__ void __launch_bounds__(256) k_ing_write(const u8 * __restrict__ text, const u64 * __restrict__ ls, u64 n_lines, int format,
                                                   const u64 * __restrict__ kept_off, const u64 * __restrict__ rec_before,
                                                   const u32 * __restrict__ hdr, int cut_id_at_space, u8 * __restrict__ bases,
                                                   u64 * __restrict__ read_off, u64 * __restrict__ id_off, u32 * __restrict__ id_len)
{
    u64 l = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31;
    if (l >= n_lines) return;
    const u64 b = ls[l], e = ls[l + 1] - 1;
    if (hdr[l])
    {
        const u64 r = rec_before[l];
        u64 ib = b + 1, ie = e;
        if (ie > ib && text[ie - 1] == '\r') ie--;
        if (cut_id_at_space)
        {
            u64 first = ie;   // first ' ' of the id
            for (u64 i = ib + lane; i < ie; i += 32) if (text[i] == ' ') { first = i; break; }
            for (int o = 16; o; o >>= 1) { u64 t2 = __shfl_xor_sync(0xffffffffu, first, o); first = min(first, t2); }
            ie = first;
        }
        if (lane == 0) { read_off[r] = kept_off[l]; id_off[r] = ib; id_len[r] = (u32)(ie - ib); }
        return;
    }
    const bool is_seq = format == 1 ? true : (l & 3) == 1;
    if (!is_seq) return;
    u64 pos = kept_off[l];
    for (u64 c0 = b; c0 < e; c0 += 32)
    {
        u64 i = c0 + lane;
        u8 ch = i < e ? text[i] : (u8)'\r';
        bool keep = ch != '\r' && (format == 2 || ch != ' ');   // the FASTQ reader drops only the carriage return
        u32 m = __ballot_sync(0xffffffffu, keep);
        if (keep) bases[pos + __popc(m & ((1u << lane) - 1))] = (u8)ing_ord5(ch);
        pos += __popc(m);
    }
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
    CKR(ctx->ing[4].reserve((n_lines + STILE + 1) * sizeof(u32)));
    CKR(ctx->ing[5].reserve((n_lines + STILE + 1) * sizeof(u32)));
    CKR(ctx->ing[6].reserve((n_lines + STILE + 1) * sizeof(u64)));
    CKR(ctx->ing[7].reserve((n_lines + STILE + 1) * sizeof(u64)));
    d_kept = ctx->ing[4].as<u32>(); d_hdr = ctx->ing[5].as<u32>(); d_koff = ctx->ing[6].as<u64>(); d_rb = ctx->ing[7].as<u64>();
    const u32 line_ctas = (u32)((n_lines * 32 + 255) / 256);
    {
        LaunchScope ls(ctx, "k_ing_line_info");
        k_ing_line_info<<<line_ctas, 256, 0, ctx->stream>>>(d_text, d_ls, n_lines, format, d_kept, d_hdr);
    }
    rc = device_scan<u64>(ctx, d_kept, n_lines, 0, d_koff, d_tot + 1, "k_ing_scan_kept");
    if (!rc) rc = device_scan<u64>(ctx, d_hdr, n_lines, 0, d_rb, d_tot + 2, "k_ing_scan_hdr");
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
        LaunchScope ls(ctx, "k_ing_write");
        k_ing_write<<<line_ctas, 256, 0, ctx->stream>>>(d_text, d_ls, n_lines, format, d_koff, d_rb, d_hdr, cut_id_at_space, R->d_bases, R->d_off,
                                                        R->d_id_off, R->d_id_len);
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
