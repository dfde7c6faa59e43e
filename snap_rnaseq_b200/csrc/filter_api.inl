// filter_api.inl -- snapb200_annotation_*, snapb200_filter_paired_batch and snapb200_rna_batch_* (include/snapb200.h; SURVEY.md
// section 8 row f3).  The kernel is filter_warp.cuh (one warp per pair) over the per-element rules of filterfmt.h, whose logic is
// verified on the host against the reference's AlignmentFilter (tests/test_filter_oracle.py).  Included at the end of snapb200.cu.
#include <condition_variable>
#include <map>
#include <thread>

#include "filter_warp.cuh"
#include "gtf_tables.h"

// Device buffers of one filter launch: grow once, reused by every later call (no cudaMalloc / cudaFree per batch).
struct FilterWorkspace {
    std::mutex lock;
    cudaStream_t stream = nullptr;
    DevBuf scratch, work, out, ev, flags;
    DevBuf in[17];  // staged inputs of snapb200_filter_paired_batch (host arrays -> HBM)
    std::vector<FltResult> h_out;
    void release()
    {
        scratch.release(); work.release(); out.release(); ev.release(); flags.release();
        for (DevBuf &b : in) b.release();
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr;
    }
};

struct RnaPool;
static void rna_pool_release(RnaPool *p);

struct snapb200_annotation {
    RnaPool *pool = nullptr;         // staging / intermediates of the rna batches in flight (created by the first snapb200_rna_batch_create)
    std::mutex pool_lock;
    int device = 0;
    int sm_count = 0;
    FltTables t;                     // device pointers, all owned by this handle
    const uint32_t *chr_rank = nullptr;
    std::vector<void *> allocs;
    std::vector<std::string> transcript_ids, chr_names;
    uint32_t n_genes = 0;
    FilterWorkspace ws[2];           // two concurrent callers of snapb200_filter_paired_batch; a third waits
    std::atomic<unsigned> ws_rr{0};
};

template <class T>
static int ann_upload(snapb200_annotation *a, const T *src, size_t count, const T **dst)
{
    void *p = nullptr;
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return set_error(SNAPB200_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    a->allocs.push_back(p);
    if (count && (e = cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice)) != cudaSuccess)
        return set_error(SNAPB200_ERR_CUDA, "annotation upload failed: %s", cudaGetErrorString(e));
    *dst = (const T *)p;
    return 0;
}
template <class T>
static int ann_upload(snapb200_annotation *a, const std::vector<T> &v, const T **dst) { return ann_upload(a, v.data(), v.size(), dst); }

extern "C" void snapb200_annotation_close(snapb200_annotation *a)
{
    if (!a) return;
    cudaSetDevice(a->device);
    for (FilterWorkspace &w : a->ws) w.release();
    if (a->pool) rna_pool_release(a->pool);
    for (void *p : a->allocs) cudaFree(p);
    delete a;
}

extern "C" int snapb200_sam_batch_rna(snapb200_annotation *ann, snapb200_index *genome, snapb200_index *transcriptome, const snapb200_sam_reads *reads0,
                                      const snapb200_sam_reads *reads1, const snapb200_sam_alignment *aln0, const snapb200_sam_alignment *aln1, int use_m,
                                      const char *read_group, char *out, uint64_t out_capacity, uint64_t *line_offsets)
{
    if (!ann || !genome || !transcriptome) return set_error(SNAPB200_ERR_ARG, "null argument");
    if (genome->device != ann->device || transcriptome->device != ann->device)
        return set_error(SNAPB200_ERR_ARG, "the annotation lives on device %d, the indices on %d and %d", ann->device, genome->device, transcriptome->device);
    if (genome->dev.n_pieces != ann->t.n_pieces || transcriptome->dev.n_pieces != ann->t.n_tpieces)
        return set_error(SNAPB200_ERR_ARG, "not the index pair the annotation was opened with (%u / %u pieces against %u / %u)", genome->dev.n_pieces,
                         transcriptome->dev.n_pieces, ann->t.n_pieces, ann->t.n_tpieces);
    return sam_batch_impl(genome, &transcriptome->dev, &ann->t, reads0, reads1, aln0, aln1, use_m, read_group, out, out_capacity, line_offsets);
}

extern "C" uint32_t snapb200_annotation_transcript_count(const snapb200_annotation *a) { return a ? (uint32_t)a->transcript_ids.size() : 0; }
extern "C" const char *snapb200_annotation_transcript_id(const snapb200_annotation *a, uint32_t i)
{
    return a && i < a->transcript_ids.size() ? a->transcript_ids[i].c_str() : nullptr;
}
extern "C" const char *snapb200_annotation_chromosome(const snapb200_annotation *a, uint32_t i)
{
    return a && i < a->chr_names.size() ? a->chr_names[i].c_str() : nullptr;
}

// The order of the map keys name + '_' + decimal(pos) between two different chromosomes is that of name + '_' alone unless one
// "name_" is a prefix of the other (then the digits take part).  ranks: position of every "name_" in strcmp order, or empty.
static std::vector<uint32_t> chromosome_ranks(const std::vector<std::string> &names)
{
    std::vector<std::string> keyed;
    for (const std::string &n : names) keyed.push_back(n + "_");
    for (size_t i = 0; i < keyed.size(); i++)
        for (size_t j = 0; j < keyed.size(); j++)
            if (i != j && keyed[j].compare(0, keyed[i].size(), keyed[i]) == 0) return std::vector<uint32_t>();  // prefix (or duplicate name)
    std::vector<uint32_t> order(keyed.size()), rank(keyed.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (uint32_t)i;
    std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return keyed[x] < keyed[y]; });
    for (size_t i = 0; i < order.size(); i++) rank[order[i]] = (uint32_t)i;
    return rank;
}

extern "C" int snapb200_annotation_open(snapb200_index *genome, snapb200_index *transcriptome, const char *gtf_path, snapb200_annotation **out)
{
    if (!genome || !transcriptome || !gtf_path || !out) return set_error(SNAPB200_ERR_ARG, "null argument");
    if (genome->device != transcriptome->device) return set_error(SNAPB200_ERR_ARG, "genome and transcriptome index are on different devices");
    GtfTables g;
    if (!gtf_load_tables(gtf_path, &g)) return set_error(SNAPB200_ERR_IO, "cannot read annotation %s", gtf_path);
    std::map<std::string, int> chr_of, transcript_of, gene_of;
    for (size_t i = 0; i < genome->piece_names.size(); i++) chr_of[genome->piece_names[i]] = (int)i;
    for (size_t i = 0; i < g.transcripts.size(); i++) transcript_of[g.transcripts[i].id] = (int)i;
    for (size_t i = 0; i < g.genes.size(); i++) gene_of[g.genes[i].id] = (int)i;
    std::vector<int32_t> tpiece_transcript, t_chr, t_gene, g_chr;
    std::vector<uint32_t> t_end, t_first(1, 0), f_type, f_start, f_end, g_start, g_end;
    for (size_t i = 0; i < transcriptome->piece_names.size(); i++) {
        std::map<std::string, int>::iterator it = transcript_of.find(transcriptome->piece_names[i]);
        if (it == transcript_of.end()) return set_error(SNAPB200_ERR_ARG, "transcriptome piece %s is not a transcript of %s", transcriptome->piece_names[i].c_str(), gtf_path);
        tpiece_transcript.push_back(it->second);
    }
    for (size_t i = 0; i < g.transcripts.size(); i++) {
        const GtfTranscriptRow &r = g.transcripts[i];
        std::map<std::string, int>::iterator c = chr_of.find(r.chr);
        if (c == chr_of.end()) return set_error(SNAPB200_ERR_ARG, "transcript %s is on %s, which the genome index does not have", r.id.c_str(), r.chr.c_str());
        t_chr.push_back(c->second);
        t_gene.push_back(gene_of[r.gene_id]);
        t_end.push_back(r.end);
        for (size_t k = 0; k < r.features.size(); k++) { f_type.push_back(r.features[k].type); f_start.push_back(r.features[k].start); f_end.push_back(r.features[k].end); }
        t_first.push_back((uint32_t)f_type.size());
    }
    for (size_t i = 0; i < g.genes.size(); i++) {
        std::map<std::string, int>::iterator c = chr_of.find(g.genes[i].chr);
        g_chr.push_back(c == chr_of.end() ? -1 : c->second);
        g_start.push_back(g.genes[i].start);
        g_end.push_back(g.genes[i].end);
    }
    std::vector<char> names;
    std::vector<uint32_t> name_off(1, 0);
    for (size_t i = 0; i < genome->piece_names.size(); i++) {
        names.insert(names.end(), genome->piece_names[i].begin(), genome->piece_names[i].end());
        name_off.push_back((uint32_t)names.size());
    }
    CUDA_TRY(cudaSetDevice(genome->device));
    snapb200_annotation *a = new snapb200_annotation();
    a->device = genome->device;
    a->sm_count = genome->sm_count;
    a->n_genes = (uint32_t)g.genes.size();
    a->chr_names = genome->piece_names;
    for (size_t i = 0; i < g.transcripts.size(); i++) a->transcript_ids.push_back(g.transcripts[i].id);
    memset(&a->t, 0, sizeof(a->t));
    // the handle keeps its own copies of the two piece tables, so closing an index first cannot leave the kernel reading freed HBM
    std::vector<uint32_t> pb(genome->dev.n_pieces), tpb(transcriptome->dev.n_pieces);
    cudaError_t e = cudaSuccess;
    if (!pb.empty()) e = cudaMemcpy(pb.data(), genome->dev.piece_begin, pb.size() * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && !tpb.empty()) e = cudaMemcpy(tpb.data(), transcriptome->dev.piece_begin, tpb.size() * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { delete a; return set_error(SNAPB200_ERR_CUDA, "reading the piece tables: %s", cudaGetErrorString(e)); }
    a->t.n_pieces = genome->dev.n_pieces;
    a->t.n_tpieces = transcriptome->dev.n_pieces;
    a->t.n_genes = (uint32_t)g.genes.size();
    a->t.gene_tree_min_stop = (!g.genes.empty() && g.genes.size() < 64) ? (int32_t)g.genes[0].start : INT32_MIN;  // flt_gene_found
    const std::vector<uint32_t> rank = chromosome_ranks(genome->piece_names);
    int rc = 0;
    if ((rc = ann_upload(a, pb, &a->t.piece_begin)) || (rc = ann_upload(a, tpb, &a->t.tpiece_begin)) ||
        (rc = ann_upload(a, names, &a->t.chr_names)) || (rc = ann_upload(a, name_off, &a->t.chr_name_off)) ||
        (rc = ann_upload(a, tpiece_transcript, &a->t.tpiece_transcript)) || (rc = ann_upload(a, t_chr, &a->t.t_chr)) ||
        (rc = ann_upload(a, t_gene, &a->t.t_gene)) || (rc = ann_upload(a, t_end, &a->t.t_end)) || (rc = ann_upload(a, t_first, &a->t.t_feat_first)) ||
        (rc = ann_upload(a, f_type, &a->t.f_type)) || (rc = ann_upload(a, f_start, &a->t.f_start)) || (rc = ann_upload(a, f_end, &a->t.f_end)) ||
        (rc = ann_upload(a, g_chr, &a->t.g_chr)) || (rc = ann_upload(a, g_start, &a->t.g_start)) || (rc = ann_upload(a, g_end, &a->t.g_end)) ||
        (!rank.empty() && (rc = ann_upload(a, rank, &a->chr_rank)))) {
        snapb200_annotation_close(a);
        return rc;
    }
    *out = a;
    return 0;
}

// ---- one launch of filter_warp_kernel over device-resident inputs ------------------------------------------------------------
static const uint32_t FW_PAIR_CAP = 8192, FW_PLOC_CAP = 8192;

// k: everything but scratch / work filled in by the caller (device pointers).  Enqueued on `st`; no synchronisation.
static int filter_launch(snapb200_annotation *a, FilterWarpArgs &k, DevBuf &scratch, DevBuf &work, cudaStream_t st)
{
    k.t = a->t;
    k.chr_rank = a->chr_rank;
    k.list_cap = k.mh + 1;
    if (const char *e = getenv("SNAPB200_FILTER_LIST_CAP")) {  // tests: a tiny scratch forces the needs_host path of the callers
        const long v = atol(e);
        if (v >= 1 && (uint32_t)v < k.list_cap) k.list_cap = (uint32_t)v;
    }
    k.pair_cap = FW_PAIR_CAP;
    k.ploc_cap = FW_PLOC_CAP;
    k.scratch_per_warp = fw_scratch_bytes(k.list_cap, k.pair_cap, k.ploc_cap);
    const uint32_t warps_per_cta = 8;
    uint32_t ctas = std::min<uint32_t>((k.n + warps_per_cta - 1) / warps_per_cta, (uint32_t)a->sm_count * 2);
    if (ctas < 1) ctas = 1;
    int rc;
    if ((rc = scratch.ensure((size_t)ctas * warps_per_cta * k.scratch_per_warp))) return rc;
    if ((rc = work.ensure(64))) return rc;
    k.scratch = scratch.as<uint8_t>();
    k.work = work.as<uint32_t>();
    CUDA_TRY(cudaMemsetAsync(work.p, 0, 64, st));
    filter_warp_kernel<<<ctas, warps_per_cta * 32, 0, st>>>(k);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

template <class T>
static int flt_stage(DevBuf &buf, const T *src, size_t count, const T **dst, cudaStream_t st)
{
    int rc = buf.ensure(std::max<size_t>(count, 1) * sizeof(T));
    if (rc) return rc;
    if (count) CUDA_TRY(cudaMemcpyAsync(buf.p, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
    *dst = buf.as<T>();
    return 0;
}

extern "C" int snapb200_filter_paired_batch(snapb200_annotation *a, const snapb200_filter_params *params, uint32_t n, const uint32_t *len0,
                                            const uint32_t *len1, const int32_t *n0, const uint32_t *loc0, const uint8_t *rc0, const int32_t *score0,
                                            const int32_t *n1, const uint32_t *loc1, const uint8_t *rc1, const int32_t *score1,
                                            const snapb200_paired_result *genome_pairs, const uint64_t *seg0, const uint32_t *ch_loc0,
                                            const uint16_t *ch_off0, const uint64_t *seg1, const uint32_t *ch_loc1, const uint16_t *ch_off1,
                                            snapb200_filter_result *results, snapb200_filter_event *events, uint8_t *needs_host)
{
    static_assert(sizeof(snapb200_filter_result) == sizeof(FltResult) && sizeof(snapb200_filter_event) == sizeof(FltEvent) && sizeof(snapb200_splice) == sizeof(FltSplice),
                  "ABI structs mirror filterfmt.h");
    if (!a || !params || (n && (!len0 || !len1 || !n0 || !loc0 || !rc0 || !score0 || !n1 || !loc1 || !rc1 || !score1 || !genome_pairs || !seg0 || !seg1 ||
                                !results || !events || !needs_host)))
        return set_error(SNAPB200_ERR_ARG, "null argument");
    if (!n) return 0;
    const uint32_t mh = params->max_hits_to_get;
    if (!mh || mh > 65535) return set_error(SNAPB200_ERR_ARG, "max_hits_to_get must be 1..65535");
    for (uint32_t i = 0; i < n; i++)
        if (n0[i] < 0 || n1[i] < 0 || (uint32_t)n0[i] > mh || (uint32_t)n1[i] > mh) return set_error(SNAPB200_ERR_ARG, "pair %u: hit count outside 0..max_hits_to_get", i);
    const size_t t0 = (size_t)seg0[2 * (size_t)n], t1 = (size_t)seg1[2 * (size_t)n];
    if ((t0 && (!ch_loc0 || !ch_off0)) || (t1 && (!ch_loc1 || !ch_off1))) return set_error(SNAPB200_ERR_ARG, "seg offsets announce seed tuples but the tuple arrays are NULL");
    CUDA_TRY(cudaSetDevice(a->device));
    int slot = -1;
    for (int q = 0; q < 2 && slot < 0; q++) if (a->ws[q].lock.try_lock()) slot = q;
    if (slot < 0) { slot = (int)(a->ws_rr.fetch_add(1) & 1); a->ws[slot].lock.lock(); }
    FilterWorkspace &w = a->ws[slot];
    struct Unlock { std::mutex &m; ~Unlock() { m.unlock(); } } unlock{w.lock};
    if (!w.stream) CUDA_TRY(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
    cudaStream_t st = w.stream;
    FilterWarpArgs k;
    memset(&k, 0, sizeof(k));
    k.prm.max_dist = params->max_dist; k.prm.max_spacing = params->max_spacing; k.prm.conf_diff = params->conf_diff; k.prm.force_spacing = (int32_t)params->force_spacing;
    k.n = n; k.mh = mh;
    const size_t rows = (size_t)n * mh;
    int rc = 0;
    if ((rc = flt_stage(w.in[0], len0, n, &k.len[0], st)) || (rc = flt_stage(w.in[1], len1, n, &k.len[1], st)) ||
        (rc = flt_stage(w.in[2], n0, n, &k.n_hits[0], st)) || (rc = flt_stage(w.in[3], n1, n, &k.n_hits[1], st)) ||
        (rc = flt_stage(w.in[4], loc0, rows, &k.loc[0], st)) || (rc = flt_stage(w.in[5], loc1, rows, &k.loc[1], st)) ||
        (rc = flt_stage(w.in[6], rc0, rows, &k.rc[0], st)) || (rc = flt_stage(w.in[7], rc1, rows, &k.rc[1], st)) ||
        (rc = flt_stage(w.in[8], score0, rows, &k.score[0], st)) || (rc = flt_stage(w.in[9], score1, rows, &k.score[1], st)) ||
        (rc = flt_stage(w.in[10], genome_pairs, n, &k.g, st)) ||
        (rc = flt_stage(w.in[11], (const unsigned long long *)seg0, 2 * (size_t)n + 1, &k.seg[0], st)) ||
        (rc = flt_stage(w.in[12], (const unsigned long long *)seg1, 2 * (size_t)n + 1, &k.seg[1], st)) ||
        (rc = flt_stage(w.in[13], ch_loc0, t0, &k.ch_loc[0], st)) || (rc = flt_stage(w.in[14], ch_off0, t0, &k.ch_off[0], st)) ||
        (rc = flt_stage(w.in[15], ch_loc1, t1, &k.ch_loc[1], st)) || (rc = flt_stage(w.in[16], ch_off1, t1, &k.ch_off[1], st)))
        return rc;
    if ((rc = w.out.ensure((size_t)n * sizeof(FltResult))) || (rc = w.ev.ensure((size_t)n * sizeof(FltEvent))) || (rc = w.flags.ensure(n))) return rc;
    CUDA_TRY(cudaMemsetAsync(w.out.p, 0, (size_t)n * sizeof(FltResult), st));
    CUDA_TRY(cudaMemsetAsync(w.ev.p, 0, (size_t)n * sizeof(FltEvent), st));
    k.out = w.out.as<FltResult>(); k.ev = w.ev.as<FltEvent>(); k.needs_host = w.flags.as<uint8_t>();
    if ((rc = filter_launch(a, k, w.scratch, w.work, st))) return rc;
    CUDA_TRY(cudaMemcpyAsync(results, w.out.p, (size_t)n * sizeof(FltResult), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(events, w.ev.p, (size_t)n * sizeof(FltEvent), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(needs_host, w.flags.p, n, cudaMemcpyDeviceToHost, st));
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return set_error(SNAPB200_ERR_CUDA, "filter_warp_kernel: %s", cudaGetErrorString(e));
    return 0;
}

// ---- the whole RNA pair loop of one batch, intermediates resident in HBM ------------------------------------------------------
// dense multi-hit rows of one mate -> CSR (counts scanned into offsets by cub, then one thread per row copies its hits)
__global__ void mh_compact_kernel(uint32_t n, uint32_t mh, const int32_t *counts, const uint32_t *off, const uint32_t *locs, const uint8_t *rcs,
                                  const int32_t *scores, uint32_t *o_loc, uint8_t *o_rc, int32_t *o_score)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = (uint32_t)counts[i], o = off[i];
    for (uint32_t k = 0; k < c; k++) {
        o_loc[o + k] = locs[(size_t)i * mh + k];
        o_rc[o + k] = rcs[(size_t)i * mh + k];
        o_score[o + k] = scores[(size_t)i * mh + k];
    }
}

struct PinBuf {  // pinned host memory that only grows
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e != cudaSuccess) return set_error(SNAPB200_ERR_CUDA, "cudaHostAlloc(%zu) failed: %s", want, cudaGetErrorString(e));
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

// Pinned staging and device-resident intermediates of one batch in flight.  A small pool per annotation (RNA_POOL of them, allocated
// on first use and kept): cudaMalloc / cudaFree / cudaHostAlloc synchronise the whole device, so nothing is allocated per batch or
// per host thread once the pool is warm; a batch borrows a set for the duration of its device work.
struct RnaResources {
    PinBuf in_off[2], in_bases[2], in_quals[2];
    // the SAM stage (snapb200_rna_batch_submit_sam): unclipped reads, ids and clipping in; text and line offsets out
    PinBuf s_off[2], s_bases[2], s_quals[2], s_front[2], s_clip[2], s_idoff[2], s_ids[2], s_rg, h_sam, h_samoff;
    DevBuf ds_off[2], ds_bases[2], ds_quals[2], ds_front[2], ds_clip[2], ds_idoff[2], ds_ids[2], ds_rg, ds_aln[2], ds_cigars, ds_lines, ds_len, ds_lineoff, ds_ctr, ds_out;
    PinBuf h_res, h_ev, h_flags, h_pairs, h_hoff[2], h_hloc[2], h_hrc[2], h_hscore[2], h_seg[2], h_cloc[2], h_coff[2], h_soff, h_splices, h_sover;
    // (the sessions' own buffers are reused by other callers between the two phases of a batch, so the intermediates live here)
    DevBuf d_hoff[2], d_hloc[2], d_hrc[2], d_hscore[2], d_seg[2], d_cnt[2], d_cloc[2], d_coff[2], d_keys[2], d_tmp, d_res, d_ev, d_flags, d_scount, d_soff, d_skind, d_sover, d_splices;
    void release()
    {
        PinBuf *pins[] = {&in_off[0], &in_off[1], &in_bases[0], &in_bases[1], &in_quals[0], &in_quals[1], &h_res, &h_ev, &h_flags, &h_pairs,
                          &h_hoff[0], &h_hoff[1], &h_hloc[0], &h_hloc[1], &h_hrc[0], &h_hrc[1], &h_hscore[0], &h_hscore[1], &h_seg[0], &h_seg[1],
                          &h_cloc[0], &h_cloc[1], &h_coff[0], &h_coff[1], &h_soff, &h_splices, &h_sover};
        for (PinBuf *q : pins) q->release();
        for (int e = 0; e < 2; e++) {
            PinBuf *sp[] = {&s_off[e], &s_bases[e], &s_quals[e], &s_front[e], &s_clip[e], &s_idoff[e], &s_ids[e]};
            for (PinBuf *q : sp) q->release();
            DevBuf *sd[] = {&ds_off[e], &ds_bases[e], &ds_quals[e], &ds_front[e], &ds_clip[e], &ds_idoff[e], &ds_ids[e], &ds_aln[e]};
            for (DevBuf *q : sd) q->release();
        }
        s_rg.release(); h_sam.release(); h_samoff.release();
        DevBuf *sd2[] = {&ds_rg, &ds_cigars, &ds_lines, &ds_len, &ds_lineoff, &ds_ctr, &ds_out};
        for (DevBuf *q : sd2) q->release();
        DevBuf *devs[] = {&d_hoff[0], &d_hoff[1], &d_hloc[0], &d_hloc[1], &d_hrc[0], &d_hrc[1], &d_hscore[0], &d_hscore[1], &d_seg[0], &d_seg[1],
                          &d_cnt[0], &d_cnt[1], &d_cloc[0], &d_cloc[1], &d_coff[0], &d_coff[1], &d_keys[0], &d_keys[1], &d_tmp, &d_res, &d_ev,
                          &d_flags, &d_scount, &d_soff, &d_skind, &d_sover, &d_splices};
        for (DevBuf *q : devs) q->release();
    }
};

#define RNA_POOL 4
struct RnaPool {
    std::mutex m;
    std::condition_variable cv;
    RnaResources *all[RNA_POOL] = {nullptr, nullptr, nullptr, nullptr};
    bool busy[RNA_POOL] = {false, false, false, false};
    RnaResources *acquire()
    {
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            for (int q = 0; q < RNA_POOL; q++)
                if (!busy[q]) { busy[q] = true; if (!all[q]) all[q] = new RnaResources(); return all[q]; }
            cv.wait(lk);
        }
    }
    void give_back(RnaResources *r)
    {
        std::lock_guard<std::mutex> lk(m);
        for (int q = 0; q < RNA_POOL; q++) if (all[q] == r) busy[q] = false;
        cv.notify_one();
    }
    void release() { for (int q = 0; q < RNA_POOL; q++) if (all[q]) { all[q]->release(); delete all[q]; all[q] = nullptr; } }
};

static void rna_pool_release(RnaPool *p) { p->release(); delete p; }

struct HostArr {  // plain host memory of the batch object (inputs copied at submit, outputs copied out of the pinned staging)
    std::vector<uint8_t> v;
    void set(const void *src, size_t bytes) { v.resize(bytes + 16); if (bytes) memcpy(v.data(), src, bytes); }
    template <class T> T *as() { return (T *)v.data(); }
};

struct snapb200_rna_batch {
    snapb200_annotation *ann = nullptr;
    snapb200_index *genome = nullptr, *transcriptome = nullptr;
    HostArr in_off[2], in_bases[2], in_quals[2];
    bool want_sam = false;  // snapb200_rna_batch_submit_sam
    int use_m = 0;
    std::string read_group;
    HostArr s_off[2], s_bases[2], s_quals[2], s_front[2], s_clip[2], s_idoff[2], s_ids[2], o_sam, o_samoff;
    uint32_t sam_max_len = 0;
    HostArr o_res, o_ev, o_flags, o_pairs, o_hoff[2], o_hloc[2], o_hrc[2], o_hscore[2], o_seg[2], o_cloc[2], o_coff[2], o_soff, o_splices, o_sover;
    snapb200_rna_params params;
    uint32_t n = 0;
    float device_ms = 0;
    // worker thread
    std::thread worker;
    std::mutex m;
    std::condition_variable cv;
    int state = 0;  // 0 idle, 1 submitted, 2 done, 3 shutting down
    int rc = 0;
    char error[512] = "";
};

static double rna_now()
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + t.tv_nsec * 1e-9;
}

// The arguments writePair gets for every pair: the filter's result after the forceSpacing rule of the run loop (PairedAligner.cpp:648-651).
// Pairs the reference's own filter has to decide are skipped.
__global__ void rna_sam_alignments_kernel(uint32_t n, const FltResult *res, const uint8_t *needs_host, int force_spacing, snapb200_sam_alignment *a0,
                                          snapb200_sam_alignment *a1)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const FltResult r = res[i];
    const bool drop = force_spacing && (r.status[0] == SNAPB200_SINGLE_HIT) != (r.status[1] == SNAPB200_SINGLE_HIT);
    snapb200_sam_alignment *out[2] = {a0 + i, a1 + i};
    for (int e = 0; e < 2; e++) {
        snapb200_sam_alignment a;
        a.location = drop ? 0xffffffffu : r.location[e];
        a.mapq = r.mapq[e];
        a.status = drop ? (uint8_t)SNAPB200_NOT_FOUND : r.status[e];
        a.direction = r.direction[e];
        a.skip = needs_host[i] ? 1 : 0;
        a.is_transcriptome = r.is_transcriptome[e];
        a.tlocation = r.tlocation[e];
        *out[e] = a;
    }
}

// The last stage of a submission made with snapb200_rna_batch_submit_sam, on the genome session's stream after the filter.
static int rna_sam_stage(snapb200_rna_batch *b, RnaResources &R, cudaStream_t stream)
{
    const uint32_t n = b->n, n_lines = 2 * n;
    int rc;
    SamArgs a;
    memset(&a, 0, sizeof(a));
    if ((rc = sam_names(b->genome, &a.names))) return rc;
    a.ix = b->genome->dev; a.tix = b->transcriptome->dev; a.tables = b->ann->t; a.rna = 1; a.cigar_stride = SAM_SPLICED_CIGAR_STRIDE;
    for (int e = 0; e < 2; e++) {
        const size_t nb = R.s_off[e].as<uint32_t>()[n], ni = R.s_idoff[e].as<uint32_t>()[n];
        if ((rc = R.ds_off[e].ensure((size_t)(n + 1) * 4)) || (rc = R.ds_bases[e].ensure(nb + 16)) || (rc = R.ds_quals[e].ensure(nb + 16)) ||
            (rc = R.ds_front[e].ensure((size_t)n * 2)) || (rc = R.ds_clip[e].ensure((size_t)n * 2)) || (rc = R.ds_idoff[e].ensure((size_t)(n + 1) * 4)) ||
            (rc = R.ds_ids[e].ensure(ni + 16)) || (rc = R.ds_aln[e].ensure((size_t)n * sizeof(snapb200_sam_alignment))))
            return rc;
        CUDA_TRY(cudaMemcpyAsync(R.ds_off[e].p, R.s_off[e].p, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, stream));
        if (nb) {
            CUDA_TRY(cudaMemcpyAsync(R.ds_bases[e].p, R.s_bases[e].p, nb, cudaMemcpyHostToDevice, stream));
            CUDA_TRY(cudaMemcpyAsync(R.ds_quals[e].p, R.s_quals[e].p, nb, cudaMemcpyHostToDevice, stream));
        }
        CUDA_TRY(cudaMemcpyAsync(R.ds_front[e].p, R.s_front[e].p, (size_t)n * 2, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(R.ds_clip[e].p, R.s_clip[e].p, (size_t)n * 2, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(R.ds_idoff[e].p, R.s_idoff[e].p, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, stream));
        if (ni) CUDA_TRY(cudaMemcpyAsync(R.ds_ids[e].p, R.s_ids[e].p, ni, cudaMemcpyHostToDevice, stream));
        a.in.rd[e].offsets = R.ds_off[e].as<uint32_t>(); a.in.rd[e].bases = R.ds_bases[e].as<uint8_t>(); a.in.rd[e].quals = R.ds_quals[e].as<uint8_t>();
        a.in.rd[e].front_clip = R.ds_front[e].as<uint16_t>(); a.in.rd[e].clipped_len = R.ds_clip[e].as<uint16_t>();
        a.in.rd[e].id_offsets = R.ds_idoff[e].as<uint32_t>(); a.in.rd[e].ids = R.ds_ids[e].as<uint8_t>();
        a.in.aln[e] = R.ds_aln[e].as<snapb200_sam_alignment>();
    }
    rna_sam_alignments_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, R.d_res.as<FltResult>(), R.d_flags.as<uint8_t>(), (int)(b->params.filter.force_spacing || b->params.paired.force_spacing),
                                                                   R.ds_aln[0].as<snapb200_sam_alignment>(), R.ds_aln[1].as<snapb200_sam_alignment>());
    CUDA_TRY(cudaGetLastError());
    const size_t rg_len = b->read_group.size();
    if ((rc = R.ds_rg.ensure(rg_len + 16)) || (rc = R.s_rg.ensure(rg_len + 16))) return rc;
    if (rg_len) {
        memcpy(R.s_rg.p, b->read_group.data(), rg_len);
        CUDA_TRY(cudaMemcpyAsync(R.ds_rg.p, R.s_rg.p, rg_len, cudaMemcpyHostToDevice, stream));
    }
    if ((rc = R.ds_cigars.ensure((size_t)n_lines * a.cigar_stride)) || (rc = R.ds_lines.ensure((size_t)n_lines * sizeof(SamLine))) ||
        (rc = R.ds_len.ensure((size_t)(n_lines + 1) * 8)) || (rc = R.ds_lineoff.ensure((size_t)(n_lines + 1) * 8)) || (rc = R.ds_ctr.ensure(sizeof(Counters))))
        return rc;
    a.n_lines = n_lines; a.in.paired = 1; a.use_m = b->use_m & SNAPB200_SAM_USE_M; a.bam = (b->use_m & SNAPB200_SAM_BAM_RECORDS) != 0;
    a.rg = R.ds_rg.as<char>(); a.rg_len = (uint32_t)rg_len;
    a.rl = std::max(32u, (b->sam_max_len + 15) & ~15u);
    a.cigars = R.ds_cigars.as<char>(); a.lines = R.ds_lines.as<SamLine>(); a.line_len = R.ds_len.as<uint64_t>(); a.line_off = R.ds_lineoff.as<uint64_t>();
    a.ctr = R.ds_ctr.as<Counters>();
    CUDA_TRY(cudaMemsetAsync(R.ds_ctr.p, 0, sizeof(Counters), stream));
    CUDA_TRY(cudaMemsetAsync(R.ds_len.as<uint64_t>() + n_lines, 0, 8, stream));
    const size_t smem = sam_warp_shared(a.rl) * WARPS_PER_CTA;
    int per_sm;
    const int grid = grid_for(sam_measure_kernel, smem, b->genome->sm_count, &per_sm);
    sam_measure_kernel<<<grid, CTA_THREADS, smem, stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, a.line_len, R.ds_lineoff.as<uint64_t>(), (int)n_lines + 1, stream);
    if ((rc = R.d_tmp.ensure(tmp_bytes))) return rc;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(R.d_tmp.p, tmp_bytes, a.line_len, R.ds_lineoff.as<uint64_t>(), (int)n_lines + 1, stream));
    if ((rc = R.h_samoff.ensure((size_t)(n_lines + 1) * 8))) return rc;
    CUDA_TRY(cudaMemcpyAsync(R.h_samoff.p, R.ds_lineoff.p, (size_t)(n_lines + 1) * 8, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    const uint64_t total = R.h_samoff.as<uint64_t>()[n_lines];
    if ((rc = R.h_sam.ensure(total + 16))) return rc;
    if (total) {
        if ((rc = R.ds_out.ensure(total + 16))) return rc;
        a.out = R.ds_out.as<char>();
        sam_write_kernel<<<(uint32_t)(((uint64_t)n_lines * SAM_WRITE_LANES + 255) / 256), 256, 0, stream>>>(a);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(R.h_sam.p, R.ds_out.p, total, cudaMemcpyDeviceToHost, stream));
    }
    return 0;
}

static int rna_run_on(snapb200_rna_batch *b, RnaResources &R)
{
    const uint32_t n = b->n;
    const bool timing = getenv("SNAPB200_RNA_TIMING") != nullptr;  // where a batch spends its time on the device side (stderr)
    double tt[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tm = rna_now(), tn;
    const double t_begin = tm;
#define RNA_MARK(i) do { tn = rna_now(); tt[i] += tn - tm; tm = tn; } while (0)
    const snapb200_rna_params &P = b->params;
    snapb200_read_batch r[2];
    for (int e = 0; e < 2; e++) { r[e].n = n; r[e].offsets = R.in_off[e].as<uint32_t>(); r[e].bases = R.in_bases[e].as<uint8_t>(); r[e].quals = R.in_quals[e].as<uint8_t>(); }
    uint32_t m0, m1;
    int rc;
    if ((rc = validate_batch(&r[0], &m0)) || (rc = validate_batch(&r[1], &m1))) return rc;
    const uint32_t mh = P.transcriptome.max_hits_to_get;
    if (!mh || mh > 65535) return set_error(SNAPB200_ERR_ARG, "transcriptome.max_hits_to_get must be 1..65535");
    CUDA_TRY(cudaSetDevice(b->genome->device));
    uint32_t total_hits[2] = {0, 0};
    // phase T: the transcriptome aligner's multi-hits of both mates (PairedAligner.cpp:584-605), compacted to CSR
    {
        BatchSlot slot(b->transcriptome);
        if ((rc = slot.open())) return rc;
        snapb200_session *s = slot.s;
        RNA_MARK(0);
        for (int e = 0; e < 2; e++) {
            if ((rc = snapb200_session_upload(s, 0, &r[e]))) return rc;
            RNA_MARK(7);
            if ((rc = snapb200_session_run_single(s, &P.transcriptome))) return rc;
            RNA_MARK(1);
            if ((rc = R.d_hoff[e].ensure((size_t)(n + 1) * 4))) return rc;
            size_t tmp_bytes = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, s->mh_counts.as<int32_t>(), R.d_hoff[e].as<uint32_t>(), (int)n + 1, s->stream);
            if ((rc = R.d_tmp.ensure(tmp_bytes))) return rc;
            // counts[n] is scratch past the last read: the scan's n-th output (the total) only needs inputs 0..n-1
            if ((rc = s->mh_counts.ensure((size_t)(n + 1) * 4))) return rc;
            CUDA_TRY(cub::DeviceScan::ExclusiveSum(R.d_tmp.p, tmp_bytes, s->mh_counts.as<int32_t>(), R.d_hoff[e].as<uint32_t>(), (int)n + 1, s->stream));
            CUDA_TRY(cudaMemcpyAsync(&total_hits[e], R.d_hoff[e].as<uint32_t>() + n, 4, cudaMemcpyDeviceToHost, s->stream));
            CUDA_TRY(cudaStreamSynchronize(s->stream));
            const size_t th = std::max<uint32_t>(total_hits[e], 1);
            if ((rc = R.d_hloc[e].ensure(th * 4)) || (rc = R.d_hrc[e].ensure(th)) || (rc = R.d_hscore[e].ensure(th * 4))) return rc;
            mh_compact_kernel<<<(n + 127) / 128, 128, 0, s->stream>>>(n, mh, s->mh_counts.as<int32_t>(), R.d_hoff[e].as<uint32_t>(), s->mh_locs.as<uint32_t>(),
                                                                      s->mh_rcs.as<uint8_t>(), s->mh_scores.as<int32_t>(), R.d_hloc[e].as<uint32_t>(),
                                                                      R.d_hrc[e].as<uint8_t>(), R.d_hscore[e].as<int32_t>());
            CUDA_TRY(cudaGetLastError());
            if ((rc = R.h_hoff[e].ensure((size_t)(n + 1) * 4)) || (rc = R.h_hloc[e].ensure(th * 4)) || (rc = R.h_hrc[e].ensure(th)) || (rc = R.h_hscore[e].ensure(th * 4))) return rc;
            CUDA_TRY(cudaMemcpyAsync(R.h_hoff[e].p, R.d_hoff[e].p, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, s->stream));
            if (total_hits[e]) {
                CUDA_TRY(cudaMemcpyAsync(R.h_hloc[e].p, R.d_hloc[e].p, (size_t)total_hits[e] * 4, cudaMemcpyDeviceToHost, s->stream));
                CUDA_TRY(cudaMemcpyAsync(R.h_hrc[e].p, R.d_hrc[e].p, (size_t)total_hits[e], cudaMemcpyDeviceToHost, s->stream));
                CUDA_TRY(cudaMemcpyAsync(R.h_hscore[e].p, R.d_hscore[e].p, (size_t)total_hits[e] * 4, cudaMemcpyDeviceToHost, s->stream));
            }
            CUDA_TRY(cudaStreamSynchronize(s->stream));  // the session's dense rows are reused by the next mate / the next caller
            RNA_MARK(2);
        }
    }
    // phase G: the genome pair, the partial aligner's seed tuples, the filter
    {
        BatchSlot slot(b->genome);
        if ((rc = slot.open())) return rc;
        snapb200_session *s = slot.s;
        RNA_MARK(3);
        if ((rc = snapb200_session_upload(s, 0, &r[0])) || (rc = snapb200_session_upload(s, 1, &r[1]))) return rc;
        RNA_MARK(8);
        if ((rc = snapb200_session_run_paired(s, &P.paired))) return rc;
        RNA_MARK(9);
        if ((rc = R.h_pairs.ensure((size_t)n * sizeof(snapb200_paired_result)))) return rc;
        rc = snapb200_session_download_paired(s, R.h_pairs.as<snapb200_paired_result>());
        if (rc) return rc;  // includes ERR_LIMIT: the reference exits there
        if (!s->host_fix.empty())  // the libm re-evaluations of download_paired, mirrored into the resident records the filter reads
            for (const MapqFix &f : s->host_fix) {
                const uint32_t pi = f.is_paired_rule ? f.index : f.index >> 1;
                CUDA_TRY(cudaMemcpyAsync(s->paired_res.as<snapb200_paired_result>() + pi, R.h_pairs.as<snapb200_paired_result>() + pi,
                                         sizeof(snapb200_paired_result), cudaMemcpyHostToDevice, s->stream));
            }
        RNA_MARK(4);
        uint64_t tuples[2] = {0, 0};
        for (int e = 0; e < 2; e++)
            if ((rc = characterize_device(b->genome, s, e, &P.partial, std::max(m0, m1), R.d_cnt[e], R.d_seg[e], R.d_keys[0], R.d_keys[1], R.d_tmp, R.d_cloc[e],
                                          R.d_coff[e], &tuples[e]))) return rc;
        RNA_MARK(5);
        FilterWarpArgs k;
        memset(&k, 0, sizeof(k));
        k.prm.max_dist = P.filter.max_dist; k.prm.max_spacing = P.filter.max_spacing; k.prm.conf_diff = P.filter.conf_diff; k.prm.force_spacing = (int32_t)P.filter.force_spacing;
        k.n = n; k.mh = mh;
        k.len_is_offsets = 1;
        for (int e = 0; e < 2; e++) {
            k.len[e] = s->offsets[e].as<uint32_t>();
            k.mh_off[e] = R.d_hoff[e].as<uint32_t>(); k.loc[e] = R.d_hloc[e].as<uint32_t>(); k.rc[e] = R.d_hrc[e].as<uint8_t>(); k.score[e] = R.d_hscore[e].as<int32_t>();
            k.seg[e] = R.d_seg[e].as<unsigned long long>(); k.ch_loc[e] = R.d_cloc[e].as<uint32_t>(); k.ch_off[e] = R.d_coff[e].as<uint16_t>();
        }
        k.g = s->paired_res.as<snapb200_paired_result>();
        if ((rc = R.d_res.ensure((size_t)n * sizeof(FltResult))) || (rc = R.d_ev.ensure((size_t)n * sizeof(FltEvent))) || (rc = R.d_flags.ensure(n))) return rc;
        CUDA_TRY(cudaMemsetAsync(R.d_res.p, 0, (size_t)n * sizeof(FltResult), s->stream));
        CUDA_TRY(cudaMemsetAsync(R.d_ev.p, 0, (size_t)n * sizeof(FltEvent), s->stream));
        k.out = R.d_res.as<FltResult>(); k.ev = R.d_ev.as<FltEvent>(); k.needs_host = R.d_flags.as<uint8_t>();
        if ((rc = filter_launch(b->ann, k, s->f_scratch, s->f_work, s->stream))) return rc;
        s->total_launches++;
        // AlignmentFilter::UnalignedRead of the reads the filter flagged, as records (count, scan, emit)
        unsigned long long n_splices = 0;
        {
            SpliceArgs sa;
            memset(&sa, 0, sizeof(sa));
            sa.t = b->ann->t; sa.n = n; sa.seed_len = b->genome->dev.seed_len;
            sa.ev = k.ev; sa.pair_needs_host = k.needs_host;
            for (int e = 0; e < 2; e++) { sa.offsets[e] = k.len[e]; sa.seg[e] = k.seg[e]; sa.ch_loc[e] = k.ch_loc[e]; sa.ch_off[e] = k.ch_off[e]; }
            sa.seg_cap = 4096;
            if (const char *e = getenv("SNAPB200_SPLICE_SEG_CAP")) {  // tests: forces splice_overflow (the caller runs UnalignedRead itself)
                const long v = atol(e);
                if (v >= 1 && v < 4096) sa.seg_cap = (uint32_t)v;
            }
            sa.scratch_per_warp = splice_scratch_bytes(sa.seg_cap);
            const uint32_t ctas = std::max<uint32_t>(1, std::min<uint32_t>((n + 7) / 8, (uint32_t)b->ann->sm_count));
            if ((rc = s->f_scratch.ensure((size_t)ctas * 8 * sa.scratch_per_warp))) return rc;
            if ((rc = R.d_scount.ensure((size_t)(n + 1) * 8)) || (rc = R.d_soff.ensure((size_t)(n + 1) * 8)) || (rc = R.d_skind.ensure(n)) || (rc = R.d_sover.ensure(n))) return rc;
            CUDA_TRY(cudaMemsetAsync(R.d_scount.p, 0, (size_t)(n + 1) * 8, s->stream));
            CUDA_TRY(cudaMemsetAsync(R.d_skind.p, 0, n, s->stream));
            CUDA_TRY(cudaMemsetAsync(R.d_sover.p, 0, n, s->stream));
            CUDA_TRY(cudaMemsetAsync(s->f_work.p, 0, 64, s->stream));
            sa.scratch = s->f_scratch.as<uint8_t>(); sa.work = s->f_work.as<uint32_t>();
            sa.counts = R.d_scount.as<unsigned long long>(); sa.kind = R.d_skind.as<uint8_t>(); sa.overflow = R.d_sover.as<uint8_t>();
            splice_kernel<false><<<ctas, 256, 0, s->stream>>>(sa);
            CUDA_TRY(cudaGetLastError());
            size_t tmp_bytes = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, R.d_scount.as<unsigned long long>(), R.d_soff.as<unsigned long long>(), (int)n + 1, s->stream);
            if ((rc = R.d_tmp.ensure(tmp_bytes))) return rc;
            CUDA_TRY(cub::DeviceScan::ExclusiveSum(R.d_tmp.p, tmp_bytes, R.d_scount.as<unsigned long long>(), R.d_soff.as<unsigned long long>(), (int)n + 1, s->stream));
            CUDA_TRY(cudaMemcpyAsync(&n_splices, R.d_soff.as<unsigned long long>() + n, 8, cudaMemcpyDeviceToHost, s->stream));
            CUDA_TRY(cudaStreamSynchronize(s->stream));
            if ((rc = R.h_soff.ensure((size_t)(n + 1) * 8)) || (rc = R.h_sover.ensure(n)) || (rc = R.h_splices.ensure(std::max<unsigned long long>(n_splices, 1) * sizeof(FltSplice)))) return rc;
            if (n_splices) {
                if ((rc = R.d_splices.ensure(n_splices * sizeof(FltSplice)))) return rc;
                CUDA_TRY(cudaMemsetAsync(s->f_work.p, 0, 64, s->stream));
                sa.counts = R.d_soff.as<unsigned long long>();
                sa.out = R.d_splices.as<FltSplice>();
                splice_kernel<true><<<ctas, 256, 0, s->stream>>>(sa);
                CUDA_TRY(cudaGetLastError());
                CUDA_TRY(cudaMemcpyAsync(R.h_splices.p, R.d_splices.p, n_splices * sizeof(FltSplice), cudaMemcpyDeviceToHost, s->stream));
            }
            CUDA_TRY(cudaMemcpyAsync(R.h_soff.p, R.d_soff.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, s->stream));
            CUDA_TRY(cudaMemcpyAsync(R.h_sover.p, R.d_sover.p, n, cudaMemcpyDeviceToHost, s->stream));
            s->total_launches += 2;
        }
        if ((rc = R.h_res.ensure((size_t)n * sizeof(FltResult))) || (rc = R.h_ev.ensure((size_t)n * sizeof(FltEvent))) || (rc = R.h_flags.ensure(n))) return rc;
        CUDA_TRY(cudaMemcpyAsync(R.h_res.p, R.d_res.p, (size_t)n * sizeof(FltResult), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaMemcpyAsync(R.h_ev.p, R.d_ev.p, (size_t)n * sizeof(FltEvent), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaMemcpyAsync(R.h_flags.p, R.d_flags.p, n, cudaMemcpyDeviceToHost, s->stream));
        for (int e = 0; e < 2; e++) {
            const size_t t = std::max<uint64_t>(tuples[e], 1);
            if ((rc = R.h_seg[e].ensure((2 * (size_t)n + 1) * 8)) || (rc = R.h_cloc[e].ensure(t * 4)) || (rc = R.h_coff[e].ensure(t * 2))) return rc;
            CUDA_TRY(cudaMemcpyAsync(R.h_seg[e].p, R.d_seg[e].p, (2 * (size_t)n + 1) * 8, cudaMemcpyDeviceToHost, s->stream));
            if (tuples[e]) {
                CUDA_TRY(cudaMemcpyAsync(R.h_cloc[e].p, R.d_cloc[e].p, tuples[e] * 4, cudaMemcpyDeviceToHost, s->stream));
                CUDA_TRY(cudaMemcpyAsync(R.h_coff[e].p, R.d_coff[e].p, tuples[e] * 2, cudaMemcpyDeviceToHost, s->stream));
            }
        }
        if (b->want_sam) {
            if ((rc = rna_sam_stage(b, R, s->stream))) return rc;
            s->total_launches += 3;
        }
        cudaError_t e2 = cudaStreamSynchronize(s->stream);
        if (e2 != cudaSuccess) return set_error(SNAPB200_ERR_CUDA, "rna batch: %s", cudaGetErrorString(e2));
        RNA_MARK(6);
    }
    if (timing)
        fprintf(stderr, "[snapb200 rna] %u pairs on device %d, %.3f s in all: wait T slot %.3f s, multi-hit uploads %.3f + kernels %.3f, compaction + download %.3f, "
                        "wait G slot %.3f, paired upload %.3f + kernels %.3f + download %.3f, CharacterizeSeeds x2 %.3f, filter + SAM + downloads %.3f\n", n,
                b->genome->device, rna_now() - t_begin, tt[0], tt[7], tt[1], tt[2], tt[3], tt[8], tt[9], tt[4], tt[5], tt[6]);
#undef RNA_MARK
    return 0;
}

// The device work of one batch: borrow a resource set, stage the inputs in pinned memory, run, copy the outputs out of the pinned staging.
static int rna_run(snapb200_rna_batch *b)
{
    {
        std::lock_guard<std::mutex> g(b->ann->pool_lock);
        if (!b->ann->pool) b->ann->pool = new RnaPool();
    }
    RnaResources *res = b->ann->pool->acquire();
    RnaResources &R = *res;
    const uint32_t n = b->n;
    int rc = 0;
    cudaSetDevice(b->genome->device);
    for (int e = 0; e < 2 && !rc; e++) {
        const size_t nb = b->in_off[e].as<uint32_t>()[n];
        if ((rc = R.in_off[e].ensure((size_t)(n + 1) * 4)) || (rc = R.in_bases[e].ensure(nb + 16)) || (rc = R.in_quals[e].ensure(nb + 16))) break;
        memcpy(R.in_off[e].p, b->in_off[e].v.data(), (size_t)(n + 1) * 4);
        memcpy(R.in_bases[e].p, b->in_bases[e].v.data(), nb);
        memcpy(R.in_quals[e].p, b->in_quals[e].v.data(), nb);
        if (b->want_sam) {
            const size_t sb = b->s_off[e].as<uint32_t>()[n], si = b->s_idoff[e].as<uint32_t>()[n];
            if ((rc = R.s_off[e].ensure((size_t)(n + 1) * 4)) || (rc = R.s_bases[e].ensure(sb + 16)) || (rc = R.s_quals[e].ensure(sb + 16)) ||
                (rc = R.s_front[e].ensure((size_t)n * 2 + 16)) || (rc = R.s_clip[e].ensure((size_t)n * 2 + 16)) || (rc = R.s_idoff[e].ensure((size_t)(n + 1) * 4)) ||
                (rc = R.s_ids[e].ensure(si + 16)))
                break;
            memcpy(R.s_off[e].p, b->s_off[e].v.data(), (size_t)(n + 1) * 4);
            memcpy(R.s_bases[e].p, b->s_bases[e].v.data(), sb);
            memcpy(R.s_quals[e].p, b->s_quals[e].v.data(), sb);
            memcpy(R.s_front[e].p, b->s_front[e].v.data(), (size_t)n * 2);
            memcpy(R.s_clip[e].p, b->s_clip[e].v.data(), (size_t)n * 2);
            memcpy(R.s_idoff[e].p, b->s_idoff[e].v.data(), (size_t)(n + 1) * 4);
            memcpy(R.s_ids[e].p, b->s_ids[e].v.data(), si);
        }
    }
    if (!rc) rc = rna_run_on(b, R);
    if (!rc) {
        b->o_res.set(R.h_res.p, (size_t)n * sizeof(FltResult)); b->o_ev.set(R.h_ev.p, (size_t)n * sizeof(FltEvent)); b->o_flags.set(R.h_flags.p, n);
        b->o_pairs.set(R.h_pairs.p, (size_t)n * sizeof(snapb200_paired_result));
        for (int e = 0; e < 2; e++) {
            const size_t th = R.h_hoff[e].as<uint32_t>()[n], tt = (size_t)R.h_seg[e].as<uint64_t>()[2 * (size_t)n];
            b->o_hoff[e].set(R.h_hoff[e].p, (size_t)(n + 1) * 4); b->o_hloc[e].set(R.h_hloc[e].p, th * 4); b->o_hrc[e].set(R.h_hrc[e].p, th); b->o_hscore[e].set(R.h_hscore[e].p, th * 4);
            b->o_seg[e].set(R.h_seg[e].p, (2 * (size_t)n + 1) * 8); b->o_cloc[e].set(R.h_cloc[e].p, tt * 4); b->o_coff[e].set(R.h_coff[e].p, tt * 2);
        }
        const size_t ns = (size_t)R.h_soff.as<uint64_t>()[n];
        b->o_soff.set(R.h_soff.p, (size_t)(n + 1) * 8); b->o_splices.set(R.h_splices.p, ns * sizeof(FltSplice)); b->o_sover.set(R.h_sover.p, n);
        if (b->want_sam) {
            b->o_samoff.set(R.h_samoff.p, (2 * (size_t)n + 1) * 8);
            b->o_sam.set(R.h_sam.p, (size_t)R.h_samoff.as<uint64_t>()[2 * (size_t)n]);
        }
    }
    b->ann->pool->give_back(res);
    return rc;
}

static void rna_worker(snapb200_rna_batch *b)
{
    std::unique_lock<std::mutex> lk(b->m);
    for (;;) {
        b->cv.wait(lk, [&] { return b->state == 1 || b->state == 3; });
        if (b->state == 3) return;
        lk.unlock();
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        const int rc = b->n ? rna_run(b) : 0;
        clock_gettime(CLOCK_MONOTONIC, &t1);
        lk.lock();
        b->rc = rc;
        b->device_ms = (float)((t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6);
        if (rc) { strncpy(b->error, g_last_error, sizeof(b->error) - 1); b->error[sizeof(b->error) - 1] = 0; }
        b->state = 2;
        b->cv.notify_all();
    }
}

extern "C" int snapb200_rna_batch_create(snapb200_annotation *a, snapb200_index *genome, snapb200_index *transcriptome, snapb200_rna_batch **out)
{
    if (!a || !genome || !transcriptome || !out) return set_error(SNAPB200_ERR_ARG, "null argument");
    if (a->device != genome->device || a->device != transcriptome->device) return set_error(SNAPB200_ERR_ARG, "annotation, genome and transcriptome index must be on one device");
    snapb200_rna_batch *b = new snapb200_rna_batch();
    b->ann = a; b->genome = genome; b->transcriptome = transcriptome;
    b->worker = std::thread(rna_worker, b);
    *out = b;
    return 0;
}

extern "C" void snapb200_rna_batch_destroy(snapb200_rna_batch *b)
{
    if (!b) return;
    {
        std::unique_lock<std::mutex> lk(b->m);
        b->cv.wait(lk, [&] { return b->state != 1; });
        b->state = 3;
        b->cv.notify_all();
    }
    b->worker.join();
    delete b;
}

static int rna_submit(snapb200_rna_batch *b, const snapb200_rna_params *params, const snapb200_read_batch *reads0, const snapb200_read_batch *reads1,
                      const snapb200_sam_reads *sam0, const snapb200_sam_reads *sam1, int use_m, const char *read_group)
{
    if (!b || !params || !reads0 || !reads1) return set_error(SNAPB200_ERR_ARG, "null argument");
    if (reads0->n != reads1->n) return set_error(SNAPB200_ERR_ARG, "mate batches differ in size");
    const bool want_sam = sam0 != nullptr;
    uint32_t sam_max_len = 0;
    if (want_sam) {
        if (!sam1 || sam0->n != reads0->n || sam1->n != reads0->n) return set_error(SNAPB200_ERR_ARG, "the SAM read batches must hold the same pairs as the read batches");
        const snapb200_sam_reads *sr[2] = {sam0, sam1};
        const snapb200_read_batch *rr[2] = {reads0, reads1};
        for (int e = 0; e < 2 && reads0->n; e++) {
            const snapb200_sam_reads *r = sr[e];
            if (!r->offsets || !r->bases || !r->quals || !r->front_clip || !r->clipped_len || !r->id_offsets || !r->ids) return set_error(SNAPB200_ERR_ARG, "null SAM read array");
            if (r->offsets[0] != 0 || r->id_offsets[0] != 0) return set_error(SNAPB200_ERR_ARG, "SAM read offsets must start at 0");
            for (uint32_t i = 0; i < r->n; i++) {
                if (r->offsets[i + 1] < r->offsets[i] || r->id_offsets[i + 1] < r->id_offsets[i]) return set_error(SNAPB200_ERR_ARG, "offsets not monotonic at read %u", i);
                const uint32_t len = r->offsets[i + 1] - r->offsets[i];
                if ((uint32_t)r->front_clip[i] + r->clipped_len[i] > len) return set_error(SNAPB200_ERR_ARG, "read %u: clipping exceeds the read", i);
                if (rr[e]->offsets && rr[e]->offsets[i + 1] - rr[e]->offsets[i] != r->clipped_len[i])
                    return set_error(SNAPB200_ERR_ARG, "read %u: clipped_len %u is not the length of the aligned read (%u)", i, r->clipped_len[i], rr[e]->offsets[i + 1] - rr[e]->offsets[i]);
                sam_max_len = std::max(sam_max_len, len);
            }
        }
        if (sam_max_len > SNAPB200_MAX_READ_LENGTH) return set_error(SNAPB200_ERR_ARG, "read of %u bases exceeds MAX_READ_LENGTH %d", sam_max_len, SNAPB200_MAX_READ_LENGTH);
    }
    std::unique_lock<std::mutex> lk(b->m);
    if (b->state == 1) return set_error(SNAPB200_ERR_ARG, "rna batch: a submitted batch has not been waited for");
    b->want_sam = want_sam;
    b->use_m = use_m;
    b->read_group = read_group ? read_group : "";
    b->sam_max_len = sam_max_len;
    if (want_sam && reads0->n) {
        const snapb200_sam_reads *sr[2] = {sam0, sam1};
        const uint32_t ns = reads0->n;
        for (int e = 0; e < 2; e++) {
            b->s_off[e].set(sr[e]->offsets, (size_t)(ns + 1) * 4);
            b->s_bases[e].set(sr[e]->bases, sr[e]->offsets[ns]);
            b->s_quals[e].set(sr[e]->quals, sr[e]->offsets[ns]);
            b->s_front[e].set(sr[e]->front_clip, (size_t)ns * 2);
            b->s_clip[e].set(sr[e]->clipped_len, (size_t)ns * 2);
            b->s_idoff[e].set(sr[e]->id_offsets, (size_t)(ns + 1) * 4);
            b->s_ids[e].set(sr[e]->ids, sr[e]->id_offsets[ns]);
        }
    }
    const snapb200_read_batch *r[2] = {reads0, reads1};
    const uint32_t n = reads0->n;
    for (int e = 0; e < 2 && n; e++) {
        if (!r[e]->offsets || !r[e]->bases || !r[e]->quals) return set_error(SNAPB200_ERR_ARG, "null read batch");
        const size_t nb = r[e]->offsets[n];
        b->in_off[e].set(r[e]->offsets, (size_t)(n + 1) * 4);
        b->in_bases[e].set(r[e]->bases, nb);
        b->in_quals[e].set(r[e]->quals, nb);
    }
    b->params = *params;
    b->n = n;
    b->state = 1;
    b->cv.notify_all();
    return 0;
}

extern "C" int snapb200_rna_batch_submit(snapb200_rna_batch *b, const snapb200_rna_params *params, const snapb200_read_batch *reads0, const snapb200_read_batch *reads1)
{
    return rna_submit(b, params, reads0, reads1, nullptr, nullptr, 0, nullptr);
}

extern "C" int snapb200_rna_batch_submit_sam(snapb200_rna_batch *b, const snapb200_rna_params *params, const snapb200_read_batch *reads0,
                                             const snapb200_read_batch *reads1, const snapb200_sam_reads *sam0, const snapb200_sam_reads *sam1, int use_m,
                                             const char *read_group)
{
    if (!sam0 || !sam1) return set_error(SNAPB200_ERR_ARG, "null SAM read batch");
    return rna_submit(b, params, reads0, reads1, sam0, sam1, use_m, read_group);
}

extern "C" int snapb200_rna_batch_wait(snapb200_rna_batch *b, snapb200_rna_view *view)
{
    if (!b || !view) return set_error(SNAPB200_ERR_ARG, "null argument");
    std::unique_lock<std::mutex> lk(b->m);
    if (b->state == 0) return set_error(SNAPB200_ERR_ARG, "rna batch: nothing was submitted");
    b->cv.wait(lk, [&] { return b->state == 2; });
    b->state = 0;
    if (b->rc) { strncpy(g_last_error, b->error, sizeof(g_last_error) - 1); return b->rc; }
    memset(view, 0, sizeof(*view));
    view->n = b->n;
    view->device_ms = b->device_ms;
    if (!b->n) return 0;
    view->results = b->o_res.as<snapb200_filter_result>();
    view->events = b->o_ev.as<snapb200_filter_event>();
    view->needs_host = b->o_flags.as<uint8_t>();
    view->genome_pairs = b->o_pairs.as<snapb200_paired_result>();
    for (int e = 0; e < 2; e++) {
        view->hit_offsets[e] = b->o_hoff[e].as<uint32_t>(); view->hit_locations[e] = b->o_hloc[e].as<uint32_t>();
        view->hit_rcs[e] = b->o_hrc[e].as<uint8_t>(); view->hit_scores[e] = b->o_hscore[e].as<int32_t>();
        view->seg_offsets[e] = b->o_seg[e].as<uint64_t>(); view->ch_locations[e] = b->o_cloc[e].as<uint32_t>(); view->ch_seed_offsets[e] = b->o_coff[e].as<uint16_t>();
    }
    view->splice_offsets = b->o_soff.as<uint64_t>();
    view->splices = b->o_splices.as<snapb200_splice>();
    view->splice_overflow = b->o_sover.as<uint8_t>();
    if (b->want_sam) {
        view->sam_text = b->o_sam.as<char>();
        view->sam_line_offsets = b->o_samoff.as<uint64_t>();
    }
    return 0;
}
