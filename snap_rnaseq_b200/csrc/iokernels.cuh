// iokernels.cuh -- the I/O edges of the path as batch kernels: FASTQ text -> read arrays,
// alignments -> SAM text.  Per-record logic is in iofmt.h (each function cites the reference lines it restates); this
// file holds the data-parallel parts: newline positions by count + scan + scatter, the byte copies, the CIGAR walk
// (lv_cigar_warp, the same routine cigar_kernel uses) and the line assembly.  Both stages are streaming kernels: their
// roof is HBM bandwidth (bytes of text in + bytes of arrays out, or the reverse).
#pragma once
#include <cub/cub.cuh>

#include "iofmt.h"

// ---- FASTQ -------------------------------------------------------------------------------------------------------
#define FQ_THREADS 256
#define FQ_BYTES_PER_THREAD 64  // four 16-byte loads; the text buffer is padded to a multiple of this

__device__ __forceinline__ uint32_t fq_newline_bits(uint32_t w)
{  // 0x80 in every byte of w that equals '\n'
    const uint32_t x = w ^ 0x0a0a0a0au;
    return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu);
}

__device__ __forceinline__ int fq_count16(const uint4 v)
{
    return __popc(fq_newline_bits(v.x)) + __popc(fq_newline_bits(v.y)) + __popc(fq_newline_bits(v.z)) + __popc(fq_newline_bits(v.w));
}

// pass A: newlines per block
__global__ void __launch_bounds__(FQ_THREADS) fq_count_kernel(const uint4 *text, uint64_t n_chunks16, uint32_t *block_counts)
{
    typedef cub::BlockReduce<int, FQ_THREADS> Reduce;
    __shared__ typename Reduce::TempStorage tmp;
    const uint64_t t = (uint64_t)blockIdx.x * FQ_THREADS + threadIdx.x;
    int c = 0;
    #pragma unroll
    for (int j = 0; j < FQ_BYTES_PER_THREAD / 16; j++) {
        const uint64_t k = t * (FQ_BYTES_PER_THREAD / 16) + j;
        if (k < n_chunks16) c += fq_count16(text[k]);
    }
    const int total = Reduce(tmp).Sum(c);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = (uint32_t)total;
}

// pass B: positions of the newlines, ascending
__global__ void __launch_bounds__(FQ_THREADS) fq_positions_kernel(const uint4 *text, uint64_t n_chunks16, const uint32_t *block_base, uint32_t *nl)
{
    typedef cub::BlockScan<int, FQ_THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    const uint64_t t = (uint64_t)blockIdx.x * FQ_THREADS + threadIdx.x;
    uint4 v[FQ_BYTES_PER_THREAD / 16];
    int c = 0;
    #pragma unroll
    for (int j = 0; j < FQ_BYTES_PER_THREAD / 16; j++) {
        const uint64_t k = t * (FQ_BYTES_PER_THREAD / 16) + j;
        v[j] = k < n_chunks16 ? text[k] : make_uint4(0, 0, 0, 0);
        c += fq_count16(v[j]);
    }
    int before;
    Scan(tmp).ExclusiveSum(c, before);
    if (c == 0) return;
    uint32_t o = block_base[blockIdx.x] + (uint32_t)before;
    const uint32_t byte0 = (uint32_t)(t * FQ_BYTES_PER_THREAD);
    #pragma unroll
    for (int j = 0; j < FQ_BYTES_PER_THREAD / 16; j++) {
        const uint32_t w[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
        #pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t bits = fq_newline_bits(w[q]);
            while (bits) {
                const int b = __ffs((int)bits) - 1;  // 7, 15, 23 or 31
                nl[o++] = byte0 + j * 16 + q * 4 + (b >> 3);
                bits &= bits - 1;
            }
        }
    }
}

struct FqArgs {
    const uint8_t *text;
    uint64_t n_bytes;
    const uint32_t *nl;
    uint32_t n_reads;
    int clipping;
    FqRecord *rec;
    uint32_t *data_len, *id_len;  // [n_reads + 1], last entry 0: scanned into offsets / id_offsets
    uint16_t *front_clip, *clipped_len;
    unsigned long long *first_error;  // min over records of (record << 8 | code)
};

__global__ void fq_record_kernel(const FqArgs a)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n_reads) return;
    const FqRecord rec = fq_record(a.text, a.n_bytes, a.nl, r, a.clipping);
    a.rec[r] = rec;
    a.data_len[r] = rec.data_len;
    a.id_len[r] = rec.id_len;
    a.front_clip[r] = rec.front_clip;
    a.clipped_len[r] = rec.clipped_len;
    if (rec.error) atomicMin(a.first_error, ((unsigned long long)r << 8) | (unsigned)rec.error);
}

struct FqCopyArgs {
    const uint8_t *text;
    uint64_t n_bytes;
    const FqRecord *rec;
    uint32_t n_reads;
    const uint32_t *offsets, *id_offsets;
    uint8_t *bases, *quals, *ids;
    const unsigned long long *first_error;  // nothing is copied once a record failed (the call returns an error)
};

// n bytes from src to dst by one warp, four at a time: words are stored at dst's alignment and assembled from the two
// aligned source words that hold them (funnel shift), so neither side needs a particular alignment.  Up to 7 bytes beyond
// src + n are read (never stored): the text buffer is padded.  UPPER: TO_UPPER_CASE on every byte (Tables.cpp:74-81).
__device__ __forceinline__ uint32_t fq_upper4(uint32_t v)
{
    const uint32_t t = v & 0x7f7f7f7fu;
    const uint32_t ge = (t + 0x1f1f1f1fu) & 0x80808080u;   // byte >= 0x61
    const uint32_t le = ~(t + 0x05050505u) & 0x80808080u;  // byte <= 0x7a
    const uint32_t m = ge & le & ~v & 0x80808080u;          // ... and below 0x80
    return v - (m >> 2);
}

template <bool UPPER>
__device__ __forceinline__ void fq_copy_bytes(uint8_t *dst, const uint8_t *src, uint32_t n, uint32_t lane)
{
    const uint32_t head = min(n, (uint32_t)((4 - ((uintptr_t)dst & 3)) & 3));
    if (lane < head) dst[lane] = UPPER ? fq_upper(src[lane]) : src[lane];
    const uint32_t nw = (n - head) >> 2;
    const uint8_t *s = src + head;
    const uint32_t *sw = (const uint32_t *)((uintptr_t)s & ~(uintptr_t)3);
    const unsigned sh = (unsigned)((uintptr_t)s & 3) * 8;
    uint32_t *dw = (uint32_t *)(dst + head);
    for (uint32_t w = lane; w < nw; w += 32) {
        const uint32_t v = __funnelshift_r(sw[w], sw[w + 1], sh);
        dw[w] = UPPER ? fq_upper4(v) : v;
    }
    const uint32_t done = head + 4 * nw;
    if (lane < n - done) dst[done + lane] = UPPER ? fq_upper(src[done + lane]) : src[done + lane];
}

// one warp per record: bases (upper-cased), qualities, id
__global__ void __launch_bounds__(256) fq_copy_kernel(const FqCopyArgs a)
{
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (r >= a.n_reads || *a.first_error != ~0ull) return;
    const FqRecord rec = a.rec[r];
    const uint32_t o = a.offsets[r], io = a.id_offsets[r];
    // a quality line shorter than the bases at the very end of the text: the reference reads on (FASTQ.cpp:241); the buffer is
    // zero-padded for 64 KiB past the text, which is what is copied then
    fq_copy_bytes<true>(a.bases + o, a.text + rec.data_start, rec.data_len, lane);
    fq_copy_bytes<false>(a.quals + o, a.text + rec.qual_start, rec.data_len, lane);
    fq_copy_bytes<false>(a.ids + io, a.text + rec.id_start, rec.id_len, lane);
}

// ---- SAM ---------------------------------------------------------------------------------------------------------
struct SamArgs {
    DevIndex ix;
    DevIndex tix;       // the transcriptome (rna != 0): the text a transcriptome alignment's CIGAR is computed against
    FltTables tables;   // the annotation (rna != 0): transcript of a transcriptome piece, its exon / intron list
    int rna;
    int bam;            // BAM records (BAMFormat::writeRead) instead of SAM lines
    uint32_t cigar_stride;  // SAM_CIGAR_STRIDE, SAM_SPLICED_CIGAR_STRIDE when rna
    SamNames names;
    SamInputs in;
    uint32_t n_lines;
    int use_m;
    const char *rg;
    uint32_t rg_len;
    uint32_t rl;  // shared-memory row for a staged read
    char *cigars;  // [n_lines][cigar_stride]
    SamLine *lines;
    uint64_t *line_len;  // [n_lines + 1], last 0: scanned into line offsets
    const uint64_t *line_off;
    char *out;
    Counters *ctr;
};

__host__ __device__ inline size_t sam_warp_shared(uint32_t rl) { return cigar_warp_shared(rl); }

// pass 1: CIGAR + the length of every line.  One warp per line.
__global__ void __launch_bounds__(CTA_THREADS) sam_measure_kernel(const SamArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    uint8_t *base = smem + sam_warp_shared(a.rl) * warp;
    int16_t *L = (int16_t *)base;
    uint8_t *P = base + lv_shared_bytes();
    uint8_t *W = P + a.rl;
    #pragma unroll 1
    for (;;) {
        const uint32_t line = fetch_work(&a.ctr->work);
        if (line >= a.n_lines) break;
        const SamWho w = sam_who(a.in, line);
        SamLine ln;
        ln.qname_len = ln.seq_len = ln.qual_len = ln.cigar_len = ln.spliced = ln.n_ops = ln.ref_len = 0;
        ln.edit_distance = -1;
        if (w.skip) {
            if (lane == 0) { a.lines[line] = ln; a.line_len[line] = 0; }
            continue;
        }
        const SamReadsDev &rd = a.in.rd[w.e];
        const uint32_t off = rd.offsets[w.i];
        const uint8_t *bases = rd.bases + off, *quals = rd.quals + off;
        const SamFields f = sam_fields(a.ix.piece_begin, (int)a.ix.n_pieces, w.me, w.has_mate, w.first_in_pair, w.mate);
        // QNAME length: /1 /2 trimming for pairs, then truncation at the first space (ReadWriter.cpp:147-161, SAM.cpp:1075-1078)
        const uint8_t *id = rd.ids + rd.id_offsets[w.i];
        uint32_t qn = rd.id_offsets[w.i + 1] - rd.id_offsets[w.i];
        if (a.in.paired) {
            const SamReadsDev &r0 = a.in.rd[0], &r1 = a.in.rd[1];
            const uint32_t l0 = r0.id_offsets[w.i + 1] - r0.id_offsets[w.i], l1 = r1.id_offsets[w.i + 1] - r1.id_offsets[w.i];
            if (sam_pair_trims_ids(r0.ids + r0.id_offsets[w.i], l0, r1.ids + r1.id_offsets[w.i], l1)) qn -= 2;
        }
        uint32_t first_space = qn;
        #pragma unroll 1
        for (uint32_t b0 = 0; b0 < qn && first_space == qn; b0 += 32) {
            const uint32_t i = b0 + lane;
            const unsigned m = __ballot_sync(FULL_MASK, i < qn && id[i] == ' ');
            if (m) first_space = b0 + __ffs((int)m) - 1;
        }
        ln.qname_len = a.bam ? qn : first_space;  // BAMFormat::writeRead copies qnameLen bytes (Bam.cpp:744)
        // SEQ / QUAL lengths as "%.*s" prints them: up to the first NUL
        uint32_t seq_len = w.me.full_len, qual_len = w.me.full_len;
        #pragma unroll 1
        for (uint32_t b0 = 0; b0 < w.me.full_len; b0 += 32) {
            const uint32_t i = b0 + lane;
            bool z = false, zq = false;
            if (i < w.me.full_len) {
                const uint8_t b = f.direction == 1 ? bases[w.me.full_len - 1 - i] : bases[i];
                z = f.direction == 1 ? !(b == 'A' || b == 'C' || b == 'G' || b == 'T' || b == 'N' || b == 'n') : b == 0;
                zq = (f.direction == 1 ? quals[w.me.full_len - 1 - i] : quals[i]) == 0;
            }
            const unsigned m = __ballot_sync(FULL_MASK, z), mq = __ballot_sync(FULL_MASK, zq);
            if (m && seq_len == w.me.full_len) seq_len = b0 + __ffs((int)m) - 1;
            if (mq && qual_len == w.me.full_len) qual_len = b0 + __ffs((int)mq) - 1;
        }
        ln.seq_len = a.bam ? w.me.full_len : seq_len;
        ln.qual_len = a.bam ? w.me.full_len : qual_len;
        // CIGAR (computeCigarString, SAM.cpp:1159-1204): the clipped read, reverse-complemented for RC, against the genome -- or, for
        // a transcriptome alignment, against the transcriptome at tlocation, with the junctions of its transcript inserted
        // (SAM.cpp:1046-1061)
        char *cig = a.cigars + (size_t)line * a.cigar_stride;
        if (f.mapped) {
            const snapb200_sam_alignment al = a.in.aln[w.e][w.i];
            const bool spliced = a.rna && al.is_transcriptome != 0;
            const DevIndex &ix = spliced ? a.tix : a.ix;
            const uint32_t len = w.me.clipped_len, loc = spliced ? al.tlocation : w.me.location;
            ln.spliced = spliced;
            if (lane == 0) cig[0] = 0;
            __syncwarp();
            int e = -1;
            if (substring_ok(ix, loc, len)) {
                const uint8_t *cb = bases + w.me.front_clip;
                #pragma unroll 1
                for (uint32_t j = lane; j < len; j += 32) P[j] = f.direction == 1 ? rc_base(cb[len - 1 - j]) : cb[j];
                stage_window(ix, loc, len, W);
                LvStr s;
                s.p = P; s.ps = 1; s.plen = (int)len;
                s.t = W + WIN_SLACK; s.ts = 1; s.tlen = (int)len;
                s.t_lo = -WIN_SLACK; s.t_hi = (int)len + WIN_SLACK;
                e = lv_cigar_warp(s, MAXK - 1, L, cig, SAM_CIGAR_STRIDE, a.use_m != 0);
                ln.edit_distance = e;  // -1 / -2: the reference prints "*" and NM:i:-1 / -2 (SAM.cpp:1196-1205)
                if (e >= 0 && lane == 0) ln.cigar_len = sam_strlen(cig, SAM_CIGAR_STRIDE);
            }
            __syncwarp();
            if (spliced) {
                // the runs move to shared memory (the LV table is dead), the leader writes the spliced string over them in HBM; a
                // CIGAR that could not be computed leaves no tokens, and the reference then prints an empty field
                char *runs = (char *)L;
                const uint32_t n_runs = __shfl_sync(FULL_MASK, e >= 0 ? ln.cigar_len : 0u, 0);
                #pragma unroll 1
                for (uint32_t j = lane; j < n_runs; j += 32) runs[j] = cig[j];
                __syncwarp();
                if (lane == 0) {
                    int n = 0;
                    uint32_t calls = 0;
                    if (e >= 0) {
                        const int piece = flt_piece_at(a.tables.tpiece_begin, (int)a.tables.n_tpieces, loc);
                        const int tr = piece >= 0 ? a.tables.tpiece_transcript[piece] : -1;
                        n = sam_splice_cigar(a.tables, tr, loc - a.tables.tpiece_begin[piece < 0 ? 0 : piece] + 1, runs, n_runs, f.clip_before, f.clip_after,
                                             cig, a.cigar_stride, &calls);
                        if (n < 0) { atomicAdd(&a.ctr->n_limit, 1u); atomicMax(&a.ctr->pad[0], line + 1); ln.spliced = 2; n = 0; }
                    }
                    ln.cigar_len = (uint32_t)n;
                    if (a.bam && ln.spliced != 2) {
                        // n_cigar_op is insertSpliceJunctions' return value, which also counts runs that printed nothing (an intron of
                        // length <= 0): the reference then emits operations it never wrote.  Such a record is left to the caller.
                        uint32_t ops, ref;
                        bam_cigar_ops(cig, ln.cigar_len, &ops, &ref, (uint8_t *)0);
                        if (ops != calls) { atomicAdd(&a.ctr->n_limit, 1u); atomicMax(&a.ctr->pad[0], line + 1); ln.spliced = 2; }
                    }
                }
                __syncwarp();
            }
        }
        if (lane == 0) {
            if (a.bam) {
                bam_count_ops(f, &ln, cig);
                if (ln.qname_len > 254) { atomicAdd(&a.ctr->n_limit, 1u); atomicMax(&a.ctr->pad[0], line + 1); ln.spliced = 2; }  // the reference exits (Bam.cpp:723)
            }
            a.lines[line] = ln;
            // spliced == 2: the record is the caller's (CIGAR too long for its slot, or one of the cases above)
            a.line_len[line] = ln.spliced == 2 ? 0 : (a.bam ? bam_record_len(ln, w.me.full_len, a.rg_len) : sam_line_len(f, ln, a.names, a.rg_len));
        }
    }
}

// pass 2: the bytes.  Half a warp per line (no warp-wide primitive is used here, and two lines per warp keep twice the loads in
// flight: the kernel waits on dependent global loads, not on bandwidth); its first lane writes the short fields before SEQ, its
// second lane the ones after QUAL, all SAM_WRITE_LANES lanes copy SEQ and QUAL.
#define SAM_WRITE_LANES 16
__global__ void __launch_bounds__(256) sam_write_kernel(const SamArgs a)
{
    const uint32_t line = (uint32_t)(((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / SAM_WRITE_LANES);
    const uint32_t lane = threadIdx.x % SAM_WRITE_LANES;
    if (line >= a.n_lines) return;
    const SamWho w = sam_who(a.in, line);
    if (w.skip) return;
    const SamLine ln = a.lines[line];
    if (ln.spliced == 2) return;
    const SamReadsDev &rd = a.in.rd[w.e];
    const uint32_t off = rd.offsets[w.i];
    const SamFields f = sam_fields(a.ix.piece_begin, (int)a.ix.n_pieces, w.me, w.has_mate, w.first_in_pair, w.mate);
    char *dst = a.out + a.line_off[line];
    const uint32_t total = (uint32_t)(a.line_off[line + 1] - a.line_off[line]);
    if (a.bam) {
        uint8_t *rec = (uint8_t *)dst;
        uint8_t *seq = rec + 36 + ln.qname_len + 1 + 4 * ln.n_ops;
        if (lane == 0) bam_put_head(rec, total, rd.ids + rd.id_offsets[w.i], f, ln, w.me.full_len, a.cigars + (size_t)line * a.cigar_stride);
        bam_put_seq_qual(seq, rd.bases + off, rd.quals + off, w.me.full_len, f.direction, lane, SAM_WRITE_LANES);
        if (lane == 1) bam_put_aux(seq + (w.me.full_len + 1) / 2 + w.me.full_len, ln, a.rg, a.rg_len);
        return;
    }
    // SEQ starts where the suffix, QUAL and SEQ end: everything after SEQ has a known length
    uint32_t tail = ln.seq_len + 1 + ln.qual_len + (a.rg_len ? 6 + a.rg_len : 0) + 10 + 6 + sam_digits_i64(ln.edit_distance) + 1;
    char *seq = dst + (total - tail);
    if (lane == 0) sam_put_prefix(dst, rd.ids + rd.id_offsets[w.i], f, ln, a.names, a.cigars + (size_t)line * a.cigar_stride);
    sam_put_seq_qual(seq, rd.bases + off, rd.quals + off, w.me.full_len, f.direction, ln, lane, SAM_WRITE_LANES);
    if (lane == 1) sam_put_suffix(seq + ln.seq_len + 1 + ln.qual_len, ln, a.rg, a.rg_len);
}
