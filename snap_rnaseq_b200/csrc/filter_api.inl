// filter_api.inl -- snapb200_annotation_open / snapb200_filter_paired_batch (include/snapb200.h).
// First version: one thread per pair around flt_filter_pair (filterfmt.h), whose logic is verified on the host against the
// reference's AlignmentFilter (tests/test_filter_oracle.py).  Included at the end of snapb200.cu.
#include "filterfmt.h"
#include "gtf_tables.h"

struct snapb200_annotation {
    snapb200_index *genome = nullptr;
    FltTables t;                     // device pointers
    std::vector<void *> allocs;
    uint32_t n_transcripts = 0, n_genes = 0;
};

struct FilterArgs {
    FltTables t;
    FltParams prm;
    uint32_t n, mh;
    const uint32_t *len[2];
    const int32_t *n_hits[2];
    const uint32_t *loc[2];
    const uint8_t *rc[2];
    const int32_t *score[2];
    const snapb200_paired_result *g;
    const uint64_t *seg[2];
    const uint32_t *ch_loc[2];
    const uint16_t *ch_off[2];
    FltResult *out;
    FltEvent *ev;
    uint8_t *needs_host;
    // scratch, one slice per thread of the grid
    FltAln *lists; uint32_t list_cap;
    FltPair *pairs; uint32_t pair_cap;
    uint32_t *ploc; uint32_t ploc_cap;
};

__global__ void __launch_bounds__(128) filter_kernel(const FilterArgs a)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    FltScratch sc;
    sc.list_cap = a.list_cap; sc.pair_cap = a.pair_cap; sc.ploc_cap = a.ploc_cap;
    sc.list[0] = a.lists + (size_t)tid * 2 * (a.list_cap + 1);
    sc.list[1] = sc.list[0] + (a.list_cap + 1);
    sc.pairs = a.pairs + (size_t)tid * a.pair_cap;
    sc.ploc[0] = a.ploc + (size_t)tid * 2 * a.ploc_cap;
    sc.ploc[1] = sc.ploc[0] + a.ploc_cap;
    for (uint32_t i = tid; i < a.n; i += nthreads) {
        FltPairInput in;
        for (int e = 0; e < 2; e++) {
            in.len[e] = a.len[e][i];
            in.n_hits[e] = a.n_hits[e][i];
            in.hit_loc[e] = a.loc[e] + (size_t)i * a.mh;
            in.hit_rc[e] = a.rc[e] + (size_t)i * a.mh;
            in.hit_score[e] = a.score[e] + (size_t)i * a.mh;
            in.g_location[e] = a.g[i].location[e]; in.g_score[e] = a.g[i].score[e]; in.g_mapq[e] = a.g[i].mapq[e];
            in.g_status[e] = a.g[i].status[e]; in.g_direction[e] = a.g[i].direction[e];
            in.ch_loc[e] = a.ch_loc[e]; in.ch_off[e] = a.ch_off[e];
            for (int k = 0; k < 3; k++) in.ch_range[e][k] = a.seg[e][2 * (size_t)i + k];
        }
        FltResult r;
        FltEvent ev;
        const int rc = flt_filter_pair(a.t, a.prm, in, sc, &r, &ev);
        a.needs_host[i] = (uint8_t)rc;
        if (rc == FLT_OK) { a.out[i] = r; a.ev[i] = ev; }
    }
}

template <class T>
static int ann_upload(snapb200_annotation *a, const std::vector<T> &v, const T **dst)
{
    void *p = nullptr;
    const size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return set_error(SNAPB200_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    a->allocs.push_back(p);
    if (!v.empty()) CUDA_TRY(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dst = (const T *)p;
    return 0;
}

extern "C" void snapb200_annotation_close(snapb200_annotation *a)
{
    if (!a) return;
    if (a->genome) cudaSetDevice(a->genome->device);
    for (void *p : a->allocs) cudaFree(p);
    delete a;
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
    a->genome = genome;
    a->n_transcripts = (uint32_t)g.transcripts.size();
    a->n_genes = (uint32_t)g.genes.size();
    memset(&a->t, 0, sizeof(a->t));
    a->t.piece_begin = genome->dev.piece_begin; a->t.n_pieces = genome->dev.n_pieces;
    a->t.tpiece_begin = transcriptome->dev.piece_begin; a->t.n_tpieces = transcriptome->dev.n_pieces;
    int rc = 0;
    if ((rc = ann_upload(a, names, &a->t.chr_names)) || (rc = ann_upload(a, name_off, &a->t.chr_name_off)) ||
        (rc = ann_upload(a, tpiece_transcript, &a->t.tpiece_transcript)) || (rc = ann_upload(a, t_chr, &a->t.t_chr)) ||
        (rc = ann_upload(a, t_gene, &a->t.t_gene)) || (rc = ann_upload(a, t_end, &a->t.t_end)) || (rc = ann_upload(a, t_first, &a->t.t_feat_first)) ||
        (rc = ann_upload(a, f_type, &a->t.f_type)) || (rc = ann_upload(a, f_start, &a->t.f_start)) || (rc = ann_upload(a, f_end, &a->t.f_end)) ||
        (rc = ann_upload(a, g_chr, &a->t.g_chr)) || (rc = ann_upload(a, g_start, &a->t.g_start)) || (rc = ann_upload(a, g_end, &a->t.g_end))) {
        snapb200_annotation_close(a);
        return rc;
    }
    *out = a;
    return 0;
}

template <class T>
static int flt_to_device(std::vector<void *> &tmp, const T *src, size_t count, const T **dst, cudaStream_t st)
{
    void *p = nullptr;
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return set_error(SNAPB200_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    tmp.push_back(p);
    if (count) CUDA_TRY(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
    *dst = (const T *)p;
    return 0;
}

extern "C" int snapb200_filter_paired_batch(snapb200_annotation *a, const snapb200_filter_params *params, uint32_t n, const uint32_t *len0,
                                            const uint32_t *len1, const int32_t *n0, const uint32_t *loc0, const uint8_t *rc0, const int32_t *score0,
                                            const int32_t *n1, const uint32_t *loc1, const uint8_t *rc1, const int32_t *score1,
                                            const snapb200_paired_result *genome_pairs, const uint64_t *seg0, const uint32_t *ch_loc0,
                                            const uint16_t *ch_off0, const uint64_t *seg1, const uint32_t *ch_loc1, const uint16_t *ch_off1,
                                            snapb200_filter_result *results, snapb200_filter_event *events, uint8_t *needs_host)
{
    static_assert(sizeof(snapb200_filter_result) == sizeof(FltResult) && sizeof(snapb200_filter_event) == sizeof(FltEvent), "ABI structs mirror filterfmt.h");
    if (!a || !params || (n && (!len0 || !len1 || !n0 || !loc0 || !rc0 || !score0 || !n1 || !loc1 || !rc1 || !score1 || !genome_pairs || !seg0 || !seg1 ||
                                !results || !events || !needs_host)))
        return set_error(SNAPB200_ERR_ARG, "null argument");
    if (!n) return 0;
    const uint32_t mh = params->max_hits_to_get;
    if (!mh) return set_error(SNAPB200_ERR_ARG, "max_hits_to_get is 0");
    for (uint32_t i = 0; i < n; i++)
        if (n0[i] < 0 || n1[i] < 0 || (uint32_t)n0[i] > mh || (uint32_t)n1[i] > mh) return set_error(SNAPB200_ERR_ARG, "pair %u: hit count outside 0..max_hits_to_get", i);
    CUDA_TRY(cudaSetDevice(a->genome->device));
    cudaStream_t st = nullptr;
    CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    std::vector<void *> tmp;
    FilterArgs k;
    memset(&k, 0, sizeof(k));
    k.t = a->t;
    k.prm.max_dist = params->max_dist; k.prm.max_spacing = params->max_spacing; k.prm.conf_diff = params->conf_diff; k.prm.force_spacing = (int32_t)params->force_spacing;
    k.n = n; k.mh = mh;
    int rc = 0;
    FltResult *d_out = nullptr;
    FltEvent *d_ev = nullptr;
    do {
        const size_t rows = (size_t)n * mh;
        if ((rc = flt_to_device(tmp, len0, n, &k.len[0], st)) || (rc = flt_to_device(tmp, len1, n, &k.len[1], st)) ||
            (rc = flt_to_device(tmp, n0, n, &k.n_hits[0], st)) || (rc = flt_to_device(tmp, n1, n, &k.n_hits[1], st)) ||
            (rc = flt_to_device(tmp, loc0, rows, &k.loc[0], st)) || (rc = flt_to_device(tmp, loc1, rows, &k.loc[1], st)) ||
            (rc = flt_to_device(tmp, rc0, rows, &k.rc[0], st)) || (rc = flt_to_device(tmp, rc1, rows, &k.rc[1], st)) ||
            (rc = flt_to_device(tmp, score0, rows, &k.score[0], st)) || (rc = flt_to_device(tmp, score1, rows, &k.score[1], st)) ||
            (rc = flt_to_device(tmp, genome_pairs, n, &k.g, st)) || (rc = flt_to_device(tmp, seg0, 2 * (size_t)n + 1, &k.seg[0], st)) ||
            (rc = flt_to_device(tmp, seg1, 2 * (size_t)n + 1, &k.seg[1], st)) || (rc = flt_to_device(tmp, ch_loc0, (size_t)seg0[2 * (size_t)n], &k.ch_loc[0], st)) ||
            (rc = flt_to_device(tmp, ch_off0, (size_t)seg0[2 * (size_t)n], &k.ch_off[0], st)) ||
            (rc = flt_to_device(tmp, ch_loc1, (size_t)seg1[2 * (size_t)n], &k.ch_loc[1], st)) ||
            (rc = flt_to_device(tmp, ch_off1, (size_t)seg1[2 * (size_t)n], &k.ch_off[1], st)))
            break;
        const uint32_t threads = 128, blocks = std::min<uint32_t>((n + threads - 1) / threads, (uint32_t)a->genome->sm_count * 4);
        const size_t nthreads = (size_t)threads * blocks;
        k.list_cap = mh + 1; k.pair_cap = 4096; k.ploc_cap = 2048;
        void *p = nullptr;
        cudaError_t e;
        if ((e = cudaMalloc(&p, nthreads * 2 * (k.list_cap + 1) * sizeof(FltAln))) != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "filter scratch: %s", cudaGetErrorString(e)); break; }
        tmp.push_back(p); k.lists = (FltAln *)p;
        if ((e = cudaMalloc(&p, nthreads * k.pair_cap * sizeof(FltPair))) != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "filter scratch: %s", cudaGetErrorString(e)); break; }
        tmp.push_back(p); k.pairs = (FltPair *)p;
        if ((e = cudaMalloc(&p, nthreads * 2 * k.ploc_cap * sizeof(uint32_t))) != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "filter scratch: %s", cudaGetErrorString(e)); break; }
        tmp.push_back(p); k.ploc = (uint32_t *)p;
        if ((e = cudaMalloc(&p, (size_t)n * sizeof(FltResult))) != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "filter results: %s", cudaGetErrorString(e)); break; }
        tmp.push_back(p); d_out = (FltResult *)p;
        if ((e = cudaMalloc(&p, (size_t)n * sizeof(FltEvent))) != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "filter events: %s", cudaGetErrorString(e)); break; }
        tmp.push_back(p); d_ev = (FltEvent *)p;
        if ((e = cudaMalloc(&p, n)) != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "filter flags: %s", cudaGetErrorString(e)); break; }
        tmp.push_back(p); k.needs_host = (uint8_t *)p;
        cudaMemsetAsync(d_out, 0, (size_t)n * sizeof(FltResult), st);
        cudaMemsetAsync(d_ev, 0, (size_t)n * sizeof(FltEvent), st);
        k.out = d_out; k.ev = d_ev;
        filter_kernel<<<blocks, threads, 0, st>>>(k);
        if ((e = cudaGetLastError()) != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "filter_kernel launch: %s", cudaGetErrorString(e)); break; }
        std::vector<FltResult> h_out(n);
        cudaMemcpyAsync(h_out.data(), d_out, (size_t)n * sizeof(FltResult), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(events, d_ev, (size_t)n * sizeof(FltEvent), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(needs_host, k.needs_host, n, cudaMemcpyDeviceToHost, st);
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "filter_kernel: %s", cudaGetErrorString(e)); break; }
        for (uint32_t i = 0; i < n; i++) {
            memset(&results[i], 0, sizeof(results[i]));
            memcpy(&results[i], &h_out[i], sizeof(FltResult));
        }
    } while (0);
    for (void *p : tmp) cudaFree(p);
    cudaStreamDestroy(st);
    return rc;
}
