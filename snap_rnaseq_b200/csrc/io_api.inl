// io_api.inl -- snapb200_fastq_parse / snapb200_sam_batch (include/snapb200.h).
// Included at the end of snapb200.cu (the library is one translation unit).  Kernels: iokernels.cuh.

struct IoScratch {  // device buffers of the calling thread, kept between calls
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    DevBuf text, block_counts, block_base, nl, rec, data_len, id_len, front, clip, err, offsets, id_offsets, bases, quals, ids, cub_tmp;
    DevBuf s_off[2], s_bases[2], s_quals[2], s_front[2], s_clip[2], s_idoff[2], s_ids[2], s_aln[2];
    DevBuf cigars, lines, line_len, line_off, out, rg;
    float fastq_ms = 0, sam_ms = 0;
    void drop()  // give the device memory back (another device is asked for, or the thread ends)
    {
        if (device < 0) return;
        DevBuf *all[] = {&text, &block_counts, &block_base, &nl, &rec, &data_len, &id_len, &front, &clip, &err, &offsets, &id_offsets, &bases,
                         &quals, &ids, &cub_tmp, &cigars, &lines, &line_len, &line_off, &out, &rg,
                         &s_off[0], &s_off[1], &s_bases[0], &s_bases[1], &s_quals[0], &s_quals[1], &s_front[0], &s_front[1],
                         &s_clip[0], &s_clip[1], &s_idoff[0], &s_idoff[1], &s_ids[0], &s_ids[1], &s_aln[0], &s_aln[1]};
        if (cudaSetDevice(device) == cudaSuccess) {  // fails harmlessly once the process is tearing the context down
            for (DevBuf *b : all) b->release();
            if (stream) cudaStreamDestroy(stream);
            for (auto &e : ev) if (e) cudaEventDestroy(e);
        }
        stream = nullptr;
        for (auto &e : ev) e = nullptr;
        device = -1;
    }
    ~IoScratch() { drop(); }
    int use(int dev)
    {
        if (device != dev) {
            drop();
            CUDA_TRY(cudaSetDevice(dev));
            CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
            for (auto &e : ev) CUDA_TRY(cudaEventCreate(&e));
            device = dev;
        } else {
            CUDA_TRY(cudaSetDevice(dev));
        }
        return 0;
    }
};
static thread_local IoScratch g_io;

extern "C" int snapb200_io_last_kernel_ms(float *fastq_ms, float *sam_ms)
{
    if (fastq_ms) *fastq_ms = g_io.fastq_ms;
    if (sam_ms) *sam_ms = g_io.sam_ms;
    return 0;
}

template <class T>
static int io_scan(IoScratch &io, const T *in, T *out, size_t n)
{
    size_t bytes = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, io.stream));
    int rc = io.cub_tmp.ensure(bytes + 16);
    if (rc) return rc;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(io.cub_tmp.p, bytes, in, out, (int)n, io.stream));
    return 0;
}

extern "C" int snapb200_fastq_record_start(const uint8_t *text, uint64_t n_bytes, uint64_t *offset)
{
    if (!offset || (n_bytes && !text)) return set_error(SNAPB200_ERR_ARG, "null argument");
    *offset = n_bytes ? fq_record_start(text, n_bytes) : 0;
    return 0;
}

extern "C" int snapb200_fastq_parse(int device, const uint8_t *text, uint64_t n_bytes, int clipping, uint32_t max_reads, uint32_t *n_reads,
                                    uint64_t *bytes_consumed, uint32_t *offsets, uint8_t *bases, uint8_t *quals, uint16_t *front_clip,
                                    uint16_t *clipped_len, uint32_t *id_offsets, uint8_t *ids)
{
    if (!n_reads || !bytes_consumed || !offsets || !id_offsets || (n_bytes && (!text || !bases || !quals || !front_clip || !clipped_len || !ids)))
        return set_error(SNAPB200_ERR_ARG, "null argument");
    if (clipping < 0 || clipping > 3) return set_error(SNAPB200_ERR_ARG, "clipping %d is not a ReadClippingType", clipping);
    if (n_bytes >= 0xfff00000ull) return set_error(SNAPB200_ERR_ARG, "FASTQ chunk of %llu bytes: at most 4 GiB - 1 MiB per call", (unsigned long long)n_bytes);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return set_error(SNAPB200_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= ndev) return set_error(SNAPB200_ERR_ARG, "device %d out of range", device);
    *n_reads = 0;
    *bytes_consumed = 0;
    offsets[0] = 0;
    id_offsets[0] = 0;
    if (!n_bytes) return 0;
    IoScratch &io = g_io;
    int rc = io.use(device);
    if (rc) return rc;
    io.fastq_ms = 0;
    // text, padded with zeros: whole 64-byte thread chunks, a readable byte after the end, and room for the quality
    // over-read of a malformed last record (the reference reads as many quality bytes as there are bases)
    const uint64_t padded = (n_bytes + 65536 + 2 * FQ_BYTES_PER_THREAD) / FQ_BYTES_PER_THREAD * FQ_BYTES_PER_THREAD;
    if ((rc = io.text.ensure(padded))) return rc;
    CUDA_TRY(cudaMemsetAsync((char *)io.text.p + n_bytes, 0, padded - n_bytes, io.stream));
    CUDA_TRY(cudaMemcpyAsync(io.text.p, text, n_bytes, cudaMemcpyHostToDevice, io.stream));
    const uint64_t scan_bytes = (n_bytes + FQ_BYTES_PER_THREAD - 1) / FQ_BYTES_PER_THREAD * FQ_BYTES_PER_THREAD;
    const uint64_t n_chunks16 = scan_bytes / 16;
    const uint64_t n_threads = scan_bytes / FQ_BYTES_PER_THREAD;
    const uint32_t nblk = (uint32_t)((n_threads + FQ_THREADS - 1) / FQ_THREADS);
    if ((rc = io.block_counts.ensure((size_t)(nblk + 1) * 4)) || (rc = io.block_base.ensure((size_t)(nblk + 1) * 4))) return rc;
    CUDA_TRY(cudaMemsetAsync((uint32_t *)io.block_counts.p + nblk, 0, 4, io.stream));
    CUDA_TRY(cudaEventRecord(io.ev[0], io.stream));
    fq_count_kernel<<<nblk, FQ_THREADS, 0, io.stream>>>(io.text.as<uint4>(), n_chunks16, io.block_counts.as<uint32_t>());
    if ((rc = io_scan(io, io.block_counts.as<uint32_t>(), io.block_base.as<uint32_t>(), (size_t)nblk + 1))) return rc;
    CUDA_TRY(cudaEventRecord(io.ev[1], io.stream));
    uint32_t n_lines = 0;
    CUDA_TRY(cudaMemcpyAsync(&n_lines, io.block_base.as<uint32_t>() + nblk, 4, cudaMemcpyDeviceToHost, io.stream));
    CUDA_TRY(cudaStreamSynchronize(io.stream));
    CUDA_TRY(cudaGetLastError());
    const uint32_t n = n_lines / 4;
    if (n > max_reads) return set_error(SNAPB200_ERR_LIMIT, "%u FASTQ records in the chunk, room for %u", n, max_reads);
    if (!n) {
        float ms = 0;
        cudaEventElapsedTime(&ms, io.ev[0], io.ev[1]);
        io.fastq_ms = ms;
        return 0;
    }
    if ((rc = io.nl.ensure((size_t)n_lines * 4)) || (rc = io.rec.ensure((size_t)n * sizeof(FqRecord))) || (rc = io.data_len.ensure((size_t)(n + 1) * 4)) ||
        (rc = io.id_len.ensure((size_t)(n + 1) * 4)) || (rc = io.front.ensure((size_t)n * 2)) || (rc = io.clip.ensure((size_t)n * 2)) ||
        (rc = io.err.ensure(8)) || (rc = io.offsets.ensure((size_t)(n + 1) * 4)) || (rc = io.id_offsets.ensure((size_t)(n + 1) * 4)) ||
        (rc = io.bases.ensure(n_bytes)) || (rc = io.quals.ensure(n_bytes)) || (rc = io.ids.ensure(n_bytes)))
        return rc;
    CUDA_TRY(cudaMemsetAsync(io.err.p, 0xff, 8, io.stream));
    CUDA_TRY(cudaMemsetAsync(io.data_len.as<uint32_t>() + n, 0, 4, io.stream));
    CUDA_TRY(cudaMemsetAsync(io.id_len.as<uint32_t>() + n, 0, 4, io.stream));
    CUDA_TRY(cudaEventRecord(io.ev[2], io.stream));
    fq_positions_kernel<<<nblk, FQ_THREADS, 0, io.stream>>>(io.text.as<uint4>(), n_chunks16, io.block_base.as<uint32_t>(), io.nl.as<uint32_t>());
    FqArgs a;
    a.text = io.text.as<uint8_t>(); a.n_bytes = n_bytes; a.nl = io.nl.as<uint32_t>(); a.n_reads = n; a.clipping = clipping;
    a.rec = io.rec.as<FqRecord>(); a.data_len = io.data_len.as<uint32_t>(); a.id_len = io.id_len.as<uint32_t>();
    a.front_clip = io.front.as<uint16_t>(); a.clipped_len = io.clip.as<uint16_t>(); a.first_error = io.err.as<unsigned long long>();
    fq_record_kernel<<<(n + 255) / 256, 256, 0, io.stream>>>(a);
    if ((rc = io_scan(io, io.data_len.as<uint32_t>(), io.offsets.as<uint32_t>(), (size_t)n + 1))) return rc;
    if ((rc = io_scan(io, io.id_len.as<uint32_t>(), io.id_offsets.as<uint32_t>(), (size_t)n + 1))) return rc;
    FqCopyArgs c;
    c.text = io.text.as<uint8_t>(); c.n_bytes = n_bytes; c.rec = io.rec.as<FqRecord>(); c.n_reads = n; c.offsets = io.offsets.as<uint32_t>();
    c.id_offsets = io.id_offsets.as<uint32_t>(); c.bases = io.bases.as<uint8_t>(); c.quals = io.quals.as<uint8_t>(); c.ids = io.ids.as<uint8_t>();
    c.first_error = io.err.as<unsigned long long>();
    fq_copy_kernel<<<(uint32_t)(((uint64_t)n * 32 + 255) / 256), 256, 0, io.stream>>>(c);
    CUDA_TRY(cudaEventRecord(io.ev[3], io.stream));
    unsigned long long first_error = ~0ull;
    FqRecord last;
    CUDA_TRY(cudaMemcpyAsync(&first_error, io.err.p, 8, cudaMemcpyDeviceToHost, io.stream));
    CUDA_TRY(cudaMemcpyAsync(&last, io.rec.as<FqRecord>() + (n - 1), sizeof(FqRecord), cudaMemcpyDeviceToHost, io.stream));
    CUDA_TRY(cudaMemcpyAsync(offsets, io.offsets.p, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, io.stream));
    CUDA_TRY(cudaMemcpyAsync(id_offsets, io.id_offsets.p, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, io.stream));
    CUDA_TRY(cudaStreamSynchronize(io.stream));
    CUDA_TRY(cudaGetLastError());
    if (first_error != ~0ull) {
        const unsigned long long r = first_error >> 8;
        const int code = (int)(first_error & 0xff);
        offsets[0] = 0;
        id_offsets[0] = 0;
        if (code == FQ_BLANK_LINE) return set_error(SNAPB200_ERR_ARG, "Syntax error in FASTQ file: blank line. (record %llu)", r);
        if (code == FQ_BAD_START) return set_error(SNAPB200_ERR_ARG, "FASTQ file has invalid starting character (record %llu)", r);
        return set_error(SNAPB200_ERR_LIMIT, "FASTQ record %llu: read longer than 65535 bases", r);
    }
    const size_t nb = offsets[n], ni = id_offsets[n];
    if (nb) {
        CUDA_TRY(cudaMemcpyAsync(bases, io.bases.p, nb, cudaMemcpyDeviceToHost, io.stream));
        CUDA_TRY(cudaMemcpyAsync(quals, io.quals.p, nb, cudaMemcpyDeviceToHost, io.stream));
    }
    if (ni) CUDA_TRY(cudaMemcpyAsync(ids, io.ids.p, ni, cudaMemcpyDeviceToHost, io.stream));
    CUDA_TRY(cudaMemcpyAsync(front_clip, io.front.p, (size_t)n * 2, cudaMemcpyDeviceToHost, io.stream));
    CUDA_TRY(cudaMemcpyAsync(clipped_len, io.clip.p, (size_t)n * 2, cudaMemcpyDeviceToHost, io.stream));
    CUDA_TRY(cudaStreamSynchronize(io.stream));
    CUDA_TRY(cudaGetLastError());
    float ms0 = 0, ms1 = 0;
    cudaEventElapsedTime(&ms0, io.ev[0], io.ev[1]);
    cudaEventElapsedTime(&ms1, io.ev[2], io.ev[3]);
    io.fastq_ms = ms0 + ms1;  // kernels only: the host round trip for the line count sits between the two timed regions
    *n_reads = n;
    *bytes_consumed = last.end;
    return 0;
}

// ---- SAM ----------------------------------------------------------------------------------------------------------------
static int sam_names(snapb200_index *x, SamNames *out)
{
    static std::mutex m;
    std::lock_guard<std::mutex> g(m);
    if (!x->sam_names_blob) {
        std::string blob;
        std::vector<uint32_t> off(x->dev.n_pieces + 1, 0);
        for (uint32_t i = 0; i < x->dev.n_pieces; i++) {
            off[i] = (uint32_t)blob.size();
            blob += i < x->piece_names.size() ? x->piece_names[i] : ("piece" + std::to_string(i));
        }
        off[x->dev.n_pieces] = (uint32_t)blob.size();
        void *db = nullptr, *dof = nullptr;
        int rc = upload(x, blob.data(), blob.size(), &db, 0, 16);
        if (rc) return rc;
        if ((rc = upload(x, off.data(), off.size() * 4, &dof))) return rc;
        x->sam_names_off = (const uint32_t *)dof;
        x->sam_names_blob = (const char *)db;
    }
    out->blob = x->sam_names_blob;
    out->off = x->sam_names_off;
    return 0;
}

static int sam_validate(const snapb200_sam_reads *r, const snapb200_sam_alignment *al, uint32_t *max_len, bool rna)
{
    if (!r) return set_error(SNAPB200_ERR_ARG, "null reads");
    if (r->n && (!r->offsets || !r->bases || !r->quals || !r->front_clip || !r->clipped_len || !r->id_offsets || !r->ids || !al))
        return set_error(SNAPB200_ERR_ARG, "null read array");
    uint32_t m = *max_len;
    for (uint32_t i = 0; i < r->n; i++) {
        if (r->offsets[i + 1] < r->offsets[i] || r->id_offsets[i + 1] < r->id_offsets[i]) return set_error(SNAPB200_ERR_ARG, "offsets not monotonic at read %u", i);
        const uint32_t len = r->offsets[i + 1] - r->offsets[i];
        if ((uint32_t)r->front_clip[i] + r->clipped_len[i] > len) return set_error(SNAPB200_ERR_ARG, "read %u: clipping exceeds the read", i);
        m = std::max(m, len);
        if (!rna && al[i].is_transcriptome && !al[i].skip)
            return set_error(SNAPB200_ERR_ARG, "read %u is a transcriptome alignment: its CIGAR needs the annotation, use snapb200_sam_batch_rna", i);
    }
    // SAMFormat::writeRead fails for reads longer than its MAX_READ buffers (SAM.cpp:1000-1001, 868-870)
    if (m > SNAPB200_MAX_READ_LENGTH) return set_error(SNAPB200_ERR_ARG, "read of %u bases exceeds MAX_READ_LENGTH %d", m, SNAPB200_MAX_READ_LENGTH);
    *max_len = m;
    return 0;
}

// tix / tables: the transcriptome and the annotation of snapb200_sam_batch_rna (filter_api.inl), NULL for genome alignments only
static int sam_batch_impl(snapb200_index *idx, const DevIndex *tix, const FltTables *tables, const snapb200_sam_reads *reads0,
                          const snapb200_sam_reads *reads1, const snapb200_sam_alignment *aln0, const snapb200_sam_alignment *aln1, int use_m,
                          const char *read_group, char *out, uint64_t out_capacity, uint64_t *line_offsets)
{
    if (!idx || !reads0 || !line_offsets) return set_error(SNAPB200_ERR_ARG, "null argument");
    const bool paired = reads1 != nullptr, rna = tables != nullptr;
    if (paired && (reads1->n != reads0->n)) return set_error(SNAPB200_ERR_ARG, "mate batches differ in size (%u vs %u)", reads0->n, reads1->n);
    uint32_t max_len = 0;
    int rc = sam_validate(reads0, aln0, &max_len, rna);
    if (rc) return rc;
    if (paired && (rc = sam_validate(reads1, aln1, &max_len, rna))) return rc;
    const uint32_t n = reads0->n;
    if (paired && n > 0x7fffffffu) return set_error(SNAPB200_ERR_ARG, "too many pairs");
    const uint32_t n_lines = paired ? 2 * n : n;
    line_offsets[0] = 0;
    if (!n_lines) return 0;
    IoScratch &io = g_io;
    if ((rc = io.use(idx->device))) return rc;
    io.sam_ms = 0;
    SamArgs a;
    memset(&a, 0, sizeof(a));
    if ((rc = sam_names(idx, &a.names))) return rc;
    a.ix = idx->dev;
    a.rna = rna;
    a.cigar_stride = rna ? SAM_SPLICED_CIGAR_STRIDE : SAM_CIGAR_STRIDE;
    if (rna) { a.tix = *tix; a.tables = *tables; }
    const snapb200_sam_reads *rs[2] = {reads0, reads1};
    const snapb200_sam_alignment *as[2] = {aln0, aln1};
    for (int e = 0; e < (paired ? 2 : 1); e++) {
        const snapb200_sam_reads *r = rs[e];
        const size_t nb = r->offsets[n] , ni = r->id_offsets[n];
        if ((rc = io.s_off[e].ensure((size_t)(n + 1) * 4)) || (rc = io.s_bases[e].ensure(nb + 16)) || (rc = io.s_quals[e].ensure(nb + 16)) ||
            (rc = io.s_front[e].ensure((size_t)n * 2)) || (rc = io.s_clip[e].ensure((size_t)n * 2)) || (rc = io.s_idoff[e].ensure((size_t)(n + 1) * 4)) ||
            (rc = io.s_ids[e].ensure(ni + 16)) || (rc = io.s_aln[e].ensure((size_t)n * sizeof(snapb200_sam_alignment))))
            return rc;
        // offsets need not start at 0: they are used as given, relative to r->bases / r->ids
        CUDA_TRY(cudaMemcpyAsync(io.s_off[e].p, r->offsets, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, io.stream));
        if (nb) {
            CUDA_TRY(cudaMemcpyAsync(io.s_bases[e].p, r->bases, nb, cudaMemcpyHostToDevice, io.stream));
            CUDA_TRY(cudaMemcpyAsync(io.s_quals[e].p, r->quals, nb, cudaMemcpyHostToDevice, io.stream));
        }
        CUDA_TRY(cudaMemcpyAsync(io.s_front[e].p, r->front_clip, (size_t)n * 2, cudaMemcpyHostToDevice, io.stream));
        CUDA_TRY(cudaMemcpyAsync(io.s_clip[e].p, r->clipped_len, (size_t)n * 2, cudaMemcpyHostToDevice, io.stream));
        CUDA_TRY(cudaMemcpyAsync(io.s_idoff[e].p, r->id_offsets, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, io.stream));
        if (ni) CUDA_TRY(cudaMemcpyAsync(io.s_ids[e].p, r->ids, ni, cudaMemcpyHostToDevice, io.stream));
        CUDA_TRY(cudaMemcpyAsync(io.s_aln[e].p, as[e], (size_t)n * sizeof(snapb200_sam_alignment), cudaMemcpyHostToDevice, io.stream));
        a.in.rd[e].offsets = io.s_off[e].as<uint32_t>(); a.in.rd[e].bases = io.s_bases[e].as<uint8_t>(); a.in.rd[e].quals = io.s_quals[e].as<uint8_t>();
        a.in.rd[e].front_clip = io.s_front[e].as<uint16_t>(); a.in.rd[e].clipped_len = io.s_clip[e].as<uint16_t>();
        a.in.rd[e].id_offsets = io.s_idoff[e].as<uint32_t>(); a.in.rd[e].ids = io.s_ids[e].as<uint8_t>();
        a.in.aln[e] = io.s_aln[e].as<snapb200_sam_alignment>();
    }
    const size_t rg_len = read_group ? strlen(read_group) : 0;
    if ((rc = io.rg.ensure(rg_len + 16))) return rc;
    if (rg_len) CUDA_TRY(cudaMemcpyAsync(io.rg.p, read_group, rg_len, cudaMemcpyHostToDevice, io.stream));
    if ((rc = io.cigars.ensure((size_t)n_lines * a.cigar_stride)) || (rc = io.lines.ensure((size_t)n_lines * sizeof(SamLine))) ||
        (rc = io.line_len.ensure((size_t)(n_lines + 1) * 8)) || (rc = io.line_off.ensure((size_t)(n_lines + 1) * 8)) || (rc = io.err.ensure(sizeof(Counters))))
        return rc;
    a.n_lines = n_lines; a.in.paired = paired; a.use_m = use_m & SNAPB200_SAM_USE_M; a.bam = (use_m & SNAPB200_SAM_BAM_RECORDS) != 0;
    a.rg = io.rg.as<char>(); a.rg_len = (uint32_t)rg_len;
    a.rl = std::max(32u, (max_len + 15) & ~15u);
    a.cigars = io.cigars.as<char>(); a.lines = io.lines.as<SamLine>(); a.line_len = io.line_len.as<uint64_t>(); a.line_off = io.line_off.as<uint64_t>();
    a.ctr = io.err.as<Counters>();
    CUDA_TRY(cudaMemsetAsync(io.err.p, 0, sizeof(Counters), io.stream));
    CUDA_TRY(cudaMemsetAsync(io.line_len.as<uint64_t>() + n_lines, 0, 8, io.stream));
    const size_t smem = sam_warp_shared(a.rl) * WARPS_PER_CTA;
    int per_sm;
    const int grid = grid_for(sam_measure_kernel, smem, idx->sm_count, &per_sm);
    CUDA_TRY(cudaEventRecord(io.ev[0], io.stream));
    sam_measure_kernel<<<grid, CTA_THREADS, smem, io.stream>>>(a);
    if ((rc = io_scan(io, io.line_len.as<uint64_t>(), io.line_off.as<uint64_t>(), (size_t)n_lines + 1))) return rc;
    CUDA_TRY(cudaEventRecord(io.ev[1], io.stream));
    CUDA_TRY(cudaMemcpyAsync(line_offsets, io.line_off.p, (size_t)(n_lines + 1) * 8, cudaMemcpyDeviceToHost, io.stream));
    CUDA_TRY(cudaStreamSynchronize(io.stream));
    CUDA_TRY(cudaGetLastError());
    float ms0 = 0, ms1 = 0;
    cudaEventElapsedTime(&ms0, io.ev[0], io.ev[1]);
    io.sam_ms = ms0;
    if (rna) {
        Counters c;
        CUDA_TRY(cudaMemcpy(&c, io.err.p, sizeof(c), cudaMemcpyDeviceToHost));
        if (c.n_limit)
            return set_error(SNAPB200_ERR_LIMIT, "%u records cannot be formatted (a spliced CIGAR of more than %d characters; BAM: a name of more than 254 bytes, a "
                             "transcript with abutting exons under the read); the last one is line %u", c.n_limit, SAM_SPLICED_CIGAR_STRIDE - 1, c.pad[0] - 1);
    } else if (use_m & SNAPB200_SAM_BAM_RECORDS) {
        Counters c;
        CUDA_TRY(cudaMemcpy(&c, io.err.p, sizeof(c), cudaMemcpyDeviceToHost));
        if (c.n_limit) return set_error(SNAPB200_ERR_LIMIT, "BAM: %u read names exceed 254 bytes (the last one on line %u)", c.n_limit, c.pad[0] - 1);
    }
    if (!out) return 0;
    const uint64_t total = line_offsets[n_lines];
    if (total > out_capacity) return set_error(SNAPB200_ERR_ARG, "SAM text needs %llu bytes, out_capacity is %llu", (unsigned long long)total, (unsigned long long)out_capacity);
    if (!total) return 0;
    if ((rc = io.out.ensure(total + 16))) return rc;
    a.out = io.out.as<char>();
    CUDA_TRY(cudaEventRecord(io.ev[2], io.stream));
    sam_write_kernel<<<(uint32_t)(((uint64_t)n_lines * SAM_WRITE_LANES + 255) / 256), 256, 0, io.stream>>>(a);
    CUDA_TRY(cudaEventRecord(io.ev[3], io.stream));
    CUDA_TRY(cudaMemcpyAsync(out, io.out.p, total, cudaMemcpyDeviceToHost, io.stream));
    CUDA_TRY(cudaStreamSynchronize(io.stream));
    CUDA_TRY(cudaGetLastError());
    cudaEventElapsedTime(&ms1, io.ev[2], io.ev[3]);
    io.sam_ms = ms0 + ms1;
    return 0;
}

extern "C" int snapb200_sam_batch(snapb200_index *idx, const snapb200_sam_reads *reads0, const snapb200_sam_reads *reads1,
                                  const snapb200_sam_alignment *aln0, const snapb200_sam_alignment *aln1, int use_m, const char *read_group,
                                  char *out, uint64_t out_capacity, uint64_t *line_offsets)
{
    return sam_batch_impl(idx, nullptr, nullptr, reads0, reads1, aln0, aln1, use_m, read_group, out, out_capacity, line_offsets);
}

// ---- BGZF (row f4b: the compressed container of BAM output; bgzf.h, bgzf_kernels.cuh) ---------------------------------------------
struct BgzfTables { const uint32_t *crc_table = nullptr, *crc_shift = nullptr; };
static int bgzf_tables(int device, BgzfTables *out)
{
    static std::mutex m;
    static BgzfTables per_device[64];
    std::lock_guard<std::mutex> g(m);
    if (device < 0 || device >= 64) return set_error(SNAPB200_ERR_ARG, "device %d out of range", device);
    BgzfTables &t = per_device[device];
    if (!t.crc_table) {
        uint32_t table[256], shift[17 * 32];
        for (uint32_t i = 0; i < 256; i++) table[i] = bgzf_crc_table_entry(i);
        bgzf_crc_shift_build(table, shift);
        void *a = nullptr, *b = nullptr;
        CUDA_TRY(cudaMalloc(&a, sizeof(table)));
        CUDA_TRY(cudaMalloc(&b, sizeof(shift)));
        CUDA_TRY(cudaMemcpy(a, table, sizeof(table), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(b, shift, sizeof(shift), cudaMemcpyHostToDevice));
        t.crc_table = (const uint32_t *)a; t.crc_shift = (const uint32_t *)b;
    }
    *out = t;
    return 0;
}

static thread_local float g_bgzf_ms = 0;
extern "C" int snapb200_bgzf_last_kernel_ms(float *ms) { if (ms) *ms = g_bgzf_ms; return 0; }

extern "C" int snapb200_bgzf_compress(int device, const uint8_t *data, uint64_t n_bytes, uint32_t chunk, uint8_t *out, uint64_t out_capacity,
                                      uint64_t *out_bytes, uint64_t *block_offsets)
{
    if ((n_bytes && !data) || !out || !out_bytes) return set_error(SNAPB200_ERR_ARG, "null argument");
    if (chunk == 0) chunk = BGZF_MAX_CHUNK;
    if (chunk > BGZF_MAX_CHUNK) return set_error(SNAPB200_ERR_ARG, "chunk %u exceeds %u bytes", chunk, BGZF_MAX_CHUNK);
    const uint64_t nb64 = n_bytes ? (n_bytes + chunk - 1) / chunk : 1;
    if (nb64 > 0x7fffffffull) return set_error(SNAPB200_ERR_ARG, "too many blocks");
    const uint32_t n_blocks = (uint32_t)nb64;
    const uint64_t bound = n_bytes + (uint64_t)n_blocks * (BGZF_HEADER + 5 + BGZF_FOOTER);
    if (out_capacity < bound) return set_error(SNAPB200_ERR_ARG, "out_capacity %llu is below the worst case %llu", (unsigned long long)out_capacity, (unsigned long long)bound);
    IoScratch &io = g_io;
    int rc;
    if ((rc = io.use(device))) return rc;
    BgzfTables tables;
    if ((rc = bgzf_tables(device, &tables))) return rc;
    DevBuf d_in, d_slots, d_sizes, d_off, d_out, d_work;
    do {
        if ((rc = d_in.ensure(n_bytes + 16)) || (rc = d_slots.ensure((size_t)n_blocks * BGZF_SLOT)) || (rc = d_sizes.ensure((size_t)(n_blocks + 1) * 8)) ||
            (rc = d_off.ensure((size_t)(n_blocks + 1) * 8)) || (rc = d_work.ensure(64)))
            break;
        cudaError_t e = cudaSuccess;
        if (n_bytes) e = cudaMemcpyAsync(d_in.p, data, n_bytes, cudaMemcpyHostToDevice, io.stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_slots.p, 0, (size_t)n_blocks * BGZF_SLOT, io.stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_sizes.p, 0, (size_t)(n_blocks + 1) * 8, io.stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_work.p, 0, 64, io.stream);
        if (e != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "bgzf: %s", cudaGetErrorString(e)); break; }
        BgzfArgs a;
        a.in = d_in.as<uint8_t>(); a.n_bytes = n_bytes; a.chunk = chunk; a.n_blocks = n_blocks; a.slots = d_slots.as<uint8_t>();
        a.sizes = d_sizes.as<unsigned long long>(); a.crc_table = tables.crc_table; a.crc_shift = tables.crc_shift; a.work = d_work.as<uint32_t>();
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        const uint32_t grid = std::min<uint32_t>((n_blocks + BGZF_WARPS - 1) / BGZF_WARPS, (uint32_t)std::max(1, sms) * 8);
        cudaEventRecord(io.ev[0], io.stream);
        bgzf_block_kernel<<<grid, BGZF_WARPS * 32, 0, io.stream>>>(a);
        if ((rc = io_scan(io, d_sizes.as<unsigned long long>(), d_off.as<unsigned long long>(), (size_t)n_blocks + 1))) break;
        std::vector<unsigned long long> off(n_blocks + 1);
        e = cudaMemcpyAsync(off.data(), d_off.p, (size_t)(n_blocks + 1) * 8, cudaMemcpyDeviceToHost, io.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(io.stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "bgzf_block_kernel: %s", cudaGetErrorString(e)); break; }
        const uint64_t total = off[n_blocks];
        if ((rc = d_out.ensure(total + 16))) break;
        bgzf_pack_kernel<<<(uint32_t)(((uint64_t)n_blocks * 32 + 255) / 256), 256, 0, io.stream>>>(d_slots.as<uint8_t>(), d_off.as<unsigned long long>(), n_blocks, d_out.as<uint8_t>());
        cudaEventRecord(io.ev[1], io.stream);
        e = cudaMemcpyAsync(out, d_out.p, total, cudaMemcpyDeviceToHost, io.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(io.stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "bgzf_pack_kernel: %s", cudaGetErrorString(e)); break; }
        cudaEventElapsedTime(&g_bgzf_ms, io.ev[0], io.ev[1]);
        *out_bytes = total;
        if (block_offsets) for (uint32_t i = 0; i <= n_blocks; i++) block_offsets[i] = off[i];
    } while (0);
    d_in.release(); d_slots.release(); d_sizes.release(); d_off.release(); d_out.release(); d_work.release();
    return rc;
}
