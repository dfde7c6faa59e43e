// filter_warp.cuh -- AlignmentFilter on the device, one WARP per pair.
//
// What the reference does per pair on the host (SNAPLib/AlignmentFilter.cpp:140-214 AddAlignment / HashAlignment, :302-739 Filter,
// :957-1037 FindPartialMatches, :1039-1059 CheckNoRC, :1061-1180 ProcessPairs; run-loop epilogue SNAPLib/PairedAligner.cpp:646-663)
// with std::map<std::string, Alignment>, four std::vector<AlignmentPair> and std::sort.  The per-element rules are the functions of
// filterfmt.h (verified on the host against the reference, tests/test_filter_oracle.py); this file is the data-parallel schedule:
//
//   A  both alignment lists are built by all lanes: one lane per multi-hit (transcript -> genome coordinates through the exon
//      table), then the de-duplication of HashAlignment and the STRING order of the map keys as a rank sort -- every entry finds the
//      winner of its key group by replaying the reference's replace rule over the entries with the same key, winners count the
//      winners with a smaller key.  Keys compare as 64-bit integers: (rank of "name_" among the chromosome names, decimal digits of
//      the position left-justified, digit count) orders exactly like strcmp on name + '_' + decimal(pos) whenever no "name_" is a
//      prefix of another (checked on the host, FilterWarpArgs::chr_rank; otherwise flt_key_compare walks the characters);
//   B  the n0 x n1 combinations are classified one per lane; the class that decides is found from four ballot-reduced counters,
//      and what std::sort would leave in pairs[0] / pairs[1] is a reduction: minimum score, how many share it, the first in loop
//      order that has it, the second smallest.  libstdc++'s sort is stable up to 16 elements (insertion sort) and pairs[0] is unique
//      when only one combination has the minimum; only a class of more than 16 combinations with a tie at the minimum is
//      materialised (ordered ballot compaction) and sorted by the leader lane with the introsort mirror of filterfmt.h;
//   C  CheckNoRC and FindPartialMatches are warp-wide "any" searches (the CharacterizeSeeds tuples come from characterize_kernel);
//   D  the leader writes the record and the event the host replays through GTFReader's public methods.
//
// A pair whose lists or classes exceed the per-warp scratch is flagged needs_host (the reference's class runs for it on the host).
#pragma once
#include "filterfmt.h"

struct FilterWarpArgs {
    FltTables t;
    const uint32_t *chr_rank;  // [n_pieces] rank of name + '_' in strcmp order; nullptr: order not static, walk the characters
    FltParams prm;
    uint32_t n;
    // read lengths: len[e][i], or offsets of a read batch (len_is_offsets): offsets[i + 1] - offsets[i]
    const uint32_t *len[2];
    int len_is_offsets;
    // transcriptome multi-hits: rows of mh entries with counts n_hits[e][i] (mh_off == nullptr), or CSR: [mh_off[e][i], mh_off[e][i+1])
    uint32_t mh;
    const uint32_t *mh_off[2];
    const int32_t *n_hits[2];
    const uint32_t *loc[2];
    const uint8_t *rc[2];
    const int32_t *score[2];
    const snapb200_paired_result *g;
    const unsigned long long *seg[2];  // CharacterizeSeeds tuples (snapb200_characterize_batch layout)
    const uint32_t *ch_loc[2];
    const uint16_t *ch_off[2];
    FltResult *out;
    FltEvent *ev;
    uint8_t *needs_host;
    uint32_t *work;  // atomic work counter (zeroed before the launch)
    // per-warp scratch
    uint8_t *scratch;
    size_t scratch_per_warp;
    uint32_t list_cap, pair_cap, ploc_cap;
};

__host__ __device__ inline size_t fw_align8(size_t v) { return (v + 7) & ~(size_t)7; }
__host__ __device__ inline size_t fw_scratch_bytes(uint32_t list_cap, uint32_t pair_cap, uint32_t ploc_cap)
{
    const size_t m = (size_t)list_cap + 1;
    return fw_align8(m * 8) + fw_align8(m) + 3 * fw_align8(m * sizeof(FltAln)) + fw_align8((size_t)pair_cap * sizeof(FltPair)) + 2 * (size_t)ploc_cap * 8;
}

struct FwScratch {
    unsigned long long *rawkey;
    uint8_t *rawwin;
    FltAln *raw, *list[2];
    FltPair *pairs;
    uint2 *ploc[2];
};

__device__ __forceinline__ FwScratch fw_scratch(const FilterWarpArgs &a, uint32_t warp)
{
    FwScratch s;
    uint8_t *p = a.scratch + (size_t)warp * a.scratch_per_warp;
    const size_t m = (size_t)a.list_cap + 1;
    s.rawkey = (unsigned long long *)p; p += fw_align8(m * 8);
    s.rawwin = p; p += fw_align8(m);
    s.raw = (FltAln *)p; p += fw_align8(m * sizeof(FltAln));
    s.list[0] = (FltAln *)p; p += fw_align8(m * sizeof(FltAln));
    s.list[1] = (FltAln *)p; p += fw_align8(m * sizeof(FltAln));
    s.pairs = (FltPair *)p; p += fw_align8((size_t)a.pair_cap * sizeof(FltPair));
    s.ploc[0] = (uint2 *)p; p += (size_t)a.ploc_cap * 8;
    s.ploc[1] = (uint2 *)p;
    return s;
}

#define FW_NO_KEY 0xffffffffffffffffull

__device__ __forceinline__ unsigned long long fw_key(const FilterWarpArgs &a, int chr, uint32_t pos)
{
    if (!a.chr_rank) return ((unsigned long long)(uint32_t)chr << 32) | pos;
    uint32_t nd = 1;
    unsigned long long p10 = 10, scale = 1000000000ull;
    while (nd < 10 && pos >= p10) { nd++; p10 *= 10; scale /= 10; }
    return ((unsigned long long)a.chr_rank[chr] << 38) | (((unsigned long long)pos * scale) << 4) | nd;
}

__device__ __forceinline__ bool fw_key_less(const FilterWarpArgs &a, const FltAln &x, unsigned long long kx, const FltAln &y, unsigned long long ky)
{
    if (a.chr_rank) return kx < ky;
    return flt_key_compare(a.t, x.chr, x.pos, y.chr, y.pos) < 0;
}

// Phase A for one end: returns the number of entries of list (in map order); *overflow when the hits do not fit the scratch.
__device__ __noinline__ uint32_t fw_build_list(const FilterWarpArgs &a, const FwScratch &s, int e, uint32_t pair, uint32_t own_len, bool *overflow)
{
    const int lane = lane_id();
    uint32_t begin, cnt;
    if (a.mh_off[e]) { begin = a.mh_off[e][pair]; cnt = a.mh_off[e][pair + 1] - begin; }
    else { begin = pair * a.mh; cnt = (uint32_t)a.n_hits[e][pair]; }
    const uint32_t m = cnt + 1;  // the genome alignment of this end is added last (PairedAligner.cpp:626-627)
    if (m > a.list_cap + 1) { *overflow = true; return 0; }
    const snapb200_paired_result g = a.g[pair];
    #pragma unroll 1
    for (uint32_t k = lane; k < m; k += 32) {
        FltAln al;
        bool ok;
        if (k < cnt) ok = flt_make_alignment(a.t, a.loc[e][begin + k], a.rc[e][begin + k] ? 1 : 0, a.score[e][begin + k], 0, true, own_len, a.prm.max_dist, &al);
        else ok = flt_make_alignment(a.t, g.location[e], g.direction[e], g.score[e], g.mapq[e], false, own_len, a.prm.max_dist, &al);
        if (ok) { s.raw[k] = al; s.rawkey[k] = fw_key(a, al.chr, al.pos); }
        else s.rawkey[k] = FW_NO_KEY;
    }
    __syncwarp();
    // the winner of every key group: HashAlignment's rule replayed over the entries with that key, in insertion order
    uint32_t n_win = 0;
    #pragma unroll 1
    for (uint32_t k = lane; k < m; k += 32) {
        const unsigned long long kk = s.rawkey[k];
        bool win = false;
        if (kk != FW_NO_KEY) {
            uint32_t w = 0xffffffffu;
            int ws = 0;
            #pragma unroll 1
            for (uint32_t j = 0; j < m; j++) {
                if (s.rawkey[j] != kk) continue;
                const int sj = s.raw[j].score;
                if (w == 0xffffffffu || sj < ws || (sj == ws && s.raw[j].is_transcriptome)) { w = j; ws = sj; }
            }
            win = w == k;
        }
        s.rawwin[k] = win ? 1 : 0;
        n_win += win ? 1 : 0;
    }
    n_win = __reduce_add_sync(FULL_MASK, n_win);
    __syncwarp();
    // map order: a winner's position is the number of winners with a smaller key
    #pragma unroll 1
    for (uint32_t k = lane; k < m; k += 32) {
        if (!s.rawwin[k]) continue;
        const FltAln me = s.raw[k];
        const unsigned long long kk = s.rawkey[k];
        uint32_t rank = 0;
        #pragma unroll 1
        for (uint32_t j = 0; j < m; j++)
            if (s.rawwin[j] && j != k && fw_key_less(a, s.raw[j], s.rawkey[j], me, kk)) rank++;
        s.list[e][rank] = me;
    }
    __syncwarp();
    return n_win;
}

__device__ __noinline__ void fw_sort_pairs(FltPair *pairs, long n) { flt_sort_pairs(pairs, n); }

// FindPartialMatches: the representative location of every distinct CharacterizeSeeds location of both reads, then any two on the
// same chromosome closer than maxSpacing.  Returns 0 / 1, or 2 when the tuples do not fit the scratch.
__device__ __noinline__ int fw_partial_match(const FilterWarpArgs &a, const FwScratch &s, uint32_t pair, uint32_t len0, uint32_t len1)
{
    const int lane = lane_id();
    uint32_t c[2] = {0, 0};
    #pragma unroll 1
    for (int e = 0; e < 2; e++) {
        const unsigned long long lo = a.seg[e][2 * (size_t)pair], mid = a.seg[e][2 * (size_t)pair + 1], hi = a.seg[e][2 * (size_t)pair + 2];
        if (hi - lo > a.ploc_cap) return 2;
        const uint32_t len = e ? len1 : len0;
        uint32_t cnt = 0;
        #pragma unroll 1
        for (unsigned long long q0 = lo; q0 < hi; q0 += 32) {
            const unsigned long long q = q0 + lane;
            bool rep = false;
            uint32_t value = 0;
            if (q < hi) {
                const uint32_t loc = a.ch_loc[e][q];
                if (q < mid) { rep = q == lo || a.ch_loc[e][q - 1] != loc; value = loc + a.ch_off[e][q]; }                  // smallest offset of the forward map
                else { rep = q + 1 == hi || a.ch_loc[e][q + 1] != loc; value = loc + (len - a.ch_off[e][q]); }              // largest offset of the RC map
            }
            const uint32_t b = __ballot_sync(FULL_MASK, rep);
            if (rep) s.ploc[e][cnt + __popc(b & ((1u << lane) - 1))] = make_uint2(value, (uint32_t)flt_piece_at(a.t.piece_begin, (int)a.t.n_pieces, value));
            cnt += __popc(b);
        }
        c[e] = cnt;
    }
    __syncwarp();
    bool found = false;
    #pragma unroll 1
    for (uint32_t u0 = 0; u0 < c[0]; u0 += 32) {
        const uint32_t u = u0 + lane;
        bool hit = false;
        if (u < c[0]) {
            const uint2 x = s.ploc[0][u];
            if ((int)x.y >= 0) {
                const int pos0 = (int)(x.x - a.t.piece_begin[x.y] + 1);
                #pragma unroll 1
                for (uint32_t v = 0; v < c[1] && !hit; v++) {
                    const uint2 y = s.ploc[1][v];
                    if (y.y != x.y) continue;
                    const int pos1 = (int)(y.x - a.t.piece_begin[y.y] + 1);
                    const uint32_t d = (uint32_t)(pos1 > pos0 ? pos1 - pos0 : pos0 - pos1);
                    hit = d < a.prm.max_spacing;
                }
            }
        }
        if (__any_sync(FULL_MASK, hit)) { found = true; break; }
    }
    __syncwarp();
    return found ? 1 : 0;
}

__device__ __forceinline__ uint32_t fw_len(const FilterWarpArgs &a, int e, uint32_t i)
{
    return a.len_is_offsets ? a.len[e][i + 1] - a.len[e][i] : a.len[e][i];
}

// One pair; every lane returns the same code (FLT_OK / FLT_SCRATCH_TOO_SMALL); lane 0 has written the record and the event.
__device__ __noinline__ int fw_filter_pair(const FilterWarpArgs &a, const FwScratch &s, uint32_t pair)
{
    const int lane = lane_id();
    const FltTables &t = a.t;
    const uint32_t len[2] = {fw_len(a, 0, pair), fw_len(a, 1, pair)};
    bool overflow = false;
    const uint32_t n0 = fw_build_list(a, s, 0, pair, len[0], &overflow);
    const uint32_t n1 = overflow ? 0 : fw_build_list(a, s, 1, pair, len[1], &overflow);
    if (overflow) return FLT_SCRATCH_TOO_SMALL;
    const FltAln *l0 = s.list[0], *l1 = s.list[1];
    const snapb200_paired_result g = a.g[pair];
    FltResult r;
    FltEvent ev;
    for (int e = 0; e < 2; e++) {
        r.location[e] = g.location[e]; r.tlocation[e] = 0; r.score[e] = g.score[e]; r.mapq[e] = g.mapq[e];
        r.status[e] = g.status[e]; r.direction[e] = g.direction[e]; r.is_transcriptome[e] = 0;
        ev.transcript[e] = -1; ev.chr[e] = 0; ev.pos_original[e] = ev.pos[e] = ev.pos_end[e] = 0;
    }
    r.aligned_as_pair = 0; r.pad = 0;
    ev.kind = FLT_EV_NONE;
    ev.unaligned = (n0 == 0 && n1 != 0) ? 1 : (n1 == 0 && n0 != 0) ? 2 : 0;
    // B: the class of every combination (the reference's loop order: its outer map holds read 1's alignments)
    const uint32_t C = n0 * n1;  // <= (list_cap + 1)^2, checked by the host wrapper to fit 32 bits
    uint32_t c_norc = 0, c_gene = 0, c_chr = 0, c_inter = 0;
    #pragma unroll 1
    for (uint32_t c = lane; c < C; c += 32) {
        const uint32_t j = c / n0, i = c - j * n0;
        const int cls = flt_classify(t, l1[j], l0[i]);
        c_norc += cls == FLT_NO_RC; c_gene += cls == FLT_INTRAGENE; c_chr += cls == FLT_INTRACHR; c_inter += cls == FLT_INTERCHR;
    }
    c_norc = __reduce_add_sync(FULL_MASK, c_norc); c_gene = __reduce_add_sync(FULL_MASK, c_gene);
    c_chr = __reduce_add_sync(FULL_MASK, c_chr); c_inter = __reduce_add_sync(FULL_MASK, c_inter);
    const int chosen = c_gene ? FLT_INTRAGENE : c_chr ? FLT_INTRACHR : c_inter ? FLT_INTERCHR : c_norc ? FLT_NO_RC : -1;
    if (chosen < 0) {
        for (int e = 0; e < 2; e++) { r.location[e] = 0; r.tlocation[e] = 0; r.score[e] = 0; r.mapq[e] = 0; r.status[e] = 0; r.direction[e] = 0; r.is_transcriptome[e] = 0; }
    } else {
        const uint32_t count = chosen == FLT_INTRAGENE ? c_gene : chosen == FLT_INTRACHR ? c_chr : chosen == FLT_INTERCHR ? c_inter : c_norc;
        // what std::sort leaves in front: minimum score, its multiplicity, the first combination (loop order) that has it, the runner-up
        uint32_t m1 = 0xffffffffu, m2 = 0xffffffffu, nmin = 0, first = 0xffffffffu;
        #pragma unroll 1
        for (uint32_t c = lane; c < C; c += 32) {
            const uint32_t j = c / n0, i = c - j * n0;
            if (flt_classify(t, l1[j], l0[i]) != chosen) continue;
            const uint32_t sc = (uint32_t)(l0[i].score + l1[j].score);
            if (sc < m1) { m2 = m1; m1 = sc; nmin = 1; first = c; }
            else if (sc == m1) nmin++;
            else if (sc < m2) m2 = sc;
        }
        const uint32_t gm1 = __reduce_min_sync(FULL_MASK, m1);
        const uint32_t gnmin = __reduce_add_sync(FULL_MASK, m1 == gm1 ? nmin : 0u);
        const uint32_t gfirst = __reduce_min_sync(FULL_MASK, m1 == gm1 ? first : 0xffffffffu);
        const uint32_t gm2 = __reduce_min_sync(FULL_MASK, m1 == gm1 ? m2 : m1);
        uint32_t c0 = gfirst, score1 = gnmin >= 2 ? gm1 : gm2;
        if (count > 16 && gnmin > 1) {  // introsort territory with a tie at the minimum: run the mirror of std::sort on the class
            if (count > a.pair_cap) return FLT_SCRATCH_TOO_SMALL;
            uint32_t base = 0;
            #pragma unroll 1
            for (uint32_t cb = 0; cb < C; cb += 32) {
                const uint32_t c = cb + lane;
                bool in = false;
                uint32_t i = 0, j = 0;
                if (c < C) { j = c / n0; i = c - j * n0; in = flt_classify(t, l1[j], l0[i]) == chosen; }
                const uint32_t b = __ballot_sync(FULL_MASK, in);
                if (in) s.pairs[base + __popc(b & ((1u << lane) - 1))] = flt_make_pair(l0[i], l1[j], i, j);
                base += __popc(b);
            }
            __syncwarp();
            if (lane == 0) fw_sort_pairs(s.pairs, (long)count);
            __syncwarp();
            const FltPair p = s.pairs[0];
            c0 = p.a2 * n0 + p.a1;
            score1 = s.pairs[1].score;
            __syncwarp();
        }
        const uint32_t j0 = c0 / n0, i0 = c0 - j0 * n0;
        FltPair pr[2];
        pr[0] = flt_make_pair(l0[i0], l1[j0], i0, j0);
        pr[1] = pr[0];
        pr[1].score = score1;
        uint32_t genome_mapq = 70;
        flt_process_pairs(t, l0, l1, pr, count, a.prm.conf_diff, &genome_mapq, &r);
        bool report = false;
        int kind = FLT_EV_NONE;
        if (chosen == FLT_INTRAGENE) {
            report = r.status[0] == 1;
            kind = FLT_EV_INCREMENT;
            r.aligned_as_pair = 1;
        } else {
            if (chosen != FLT_NO_RC && r.status[0] == 1 && c_norc) {  // CheckNoRC
                const uint32_t sum = (uint32_t)(r.score[0] + r.score[1]);
                bool hit = false;
                #pragma unroll 1
                for (uint32_t cb = 0; cb < C && !hit; cb += 32) {
                    const uint32_t c = cb + lane;
                    bool h = false;
                    if (c < C) {
                        const uint32_t j = c / n0, i = c - j * n0;
                        h = flt_classify(t, l1[j], l0[i]) == FLT_NO_RC && l0[i].chr == l1[j].chr && (uint32_t)(l0[i].score + l1[j].score) < sum;
                    }
                    hit = __any_sync(FULL_MASK, h);
                }
                if (hit) { r.status[0] = r.status[1] = 2; r.mapq[0] = r.mapq[1] = 1; }
            }
            const bool near = chosen == FLT_INTRACHR && (uint32_t)pr[0].distance <= a.prm.max_spacing;
            if (!near) {
                if (r.status[0] == 1) {
                    const int pm = fw_partial_match(a, s, pair, len[0], len[1]);
                    if (pm == 2) return FLT_SCRATCH_TOO_SMALL;
                    if (pm == 1) { r.status[0] = r.status[1] = 2; r.mapq[0] = r.mapq[1] = 1; }
                }
                report = r.status[0] == 1;
                kind = chosen == FLT_INTRACHR ? FLT_EV_INTRACHR : chosen == FLT_INTERCHR ? FLT_EV_INTERCHR
                     : (l0[i0].chr == l1[j0].chr ? FLT_EV_INTRACHR : FLT_EV_INTERCHR);
            }
        }
        if (report) {
            ev.kind = kind;
            const FltAln *al[2] = {&l0[i0], &l1[j0]};
            for (int e = 0; e < 2; e++) {
                ev.transcript[e] = al[e]->transcript; ev.chr[e] = al[e]->chr;
                ev.pos_original[e] = al[e]->pos_original; ev.pos[e] = al[e]->pos; ev.pos_end[e] = al[e]->pos_end;
            }
        }
    }
    // the run loop's epilogue (PairedAligner.cpp:646-663)
    if (a.prm.force_spacing && (r.status[0] == 1) != (r.status[1] == 1)) { r.status[0] = r.status[1] = 0; r.location[0] = r.location[1] = FLT_INVALID_LOC; }
    if (r.score[0] + r.score[1] >= 5) {
        if (r.mapq[0] < 50) r.mapq[0] /= 2;
        if (r.mapq[1] < 50) r.mapq[1] /= 2;
    }
    for (int e = 0; e < 2; e++) if (!r.is_transcriptome[e]) r.tlocation[e] = 0;
    if (lane == 0) { a.out[pair] = r; a.ev[pair] = ev; }
    return FLT_OK;
}

__global__ void __launch_bounds__(256) filter_warp_kernel(const FilterWarpArgs a)
{
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const FwScratch s = fw_scratch(a, warp);
    for (;;) {
        const uint32_t i = fetch_work(a.work);
        if (i >= a.n) break;
        const int rc = fw_filter_pair(a, s, i);
        if (lane_id() == 0) a.needs_host[i] = (uint8_t)rc;
        __syncwarp();
    }
}

// ---- AlignmentFilter::UnalignedRead as records, one warp per flagged read (the serial specification is flt_unaligned_splices) ---------
// Pass COUNT sizes every read's record list (and says which GTFReader call it is for), a scan turns the sizes into offsets, pass EMIT
// writes the records in the reference's loop order.  Per read: the partial alignments are built by ordered ballot compaction over the
// seed tuples; every alignment gets the (at most SPLICE_GENES) genes that cover it, found by all lanes over the gene table, so that
// the gene test of a candidate is a few boundary checks; the n (n - 1) / 2 candidates are tested 32 per step.
#define SPLICE_GENES 4

struct SpliceArgs {
    FltTables t;
    uint32_t n, seed_len;
    const uint32_t *offsets[2];          // read batch offsets (read length = offsets[i + 1] - offsets[i])
    const FltEvent *ev;
    const uint8_t *pair_needs_host;      // pairs whose filter result is not valid are skipped (the host runs the reference for them)
    const unsigned long long *seg[2];
    const uint32_t *ch_loc[2];
    const uint16_t *ch_off[2];
    unsigned long long *counts;          // COUNT: [n + 1] out; EMIT: exclusive offsets in
    uint8_t *kind;                       // [n] FLT_SPLICE_* of the read's records (COUNT out, EMIT in)
    uint8_t *overflow;                   // [n] more partial alignments than the scratch holds: the host runs UnalignedRead itself
    FltSplice *out;                      // EMIT
    uint32_t *work;
    uint8_t *scratch;
    size_t scratch_per_warp;
    uint32_t seg_cap;
};

struct SpliceGenes { uint32_t n; int32_t g[SPLICE_GENES]; };  // n > SPLICE_GENES: too many, scan the table

__host__ __device__ inline size_t splice_scratch_bytes(uint32_t seg_cap) { return (size_t)seg_cap * (sizeof(FltSeg) + sizeof(SpliceGenes)); }

__device__ __forceinline__ bool splice_in_gene(const FltTables &t, const FltSeg &a0, const SpliceGenes &sg, const FltSeg &a1)
{
    if (sg.n > SPLICE_GENES) return flt_splice_in_gene(t, a0, a1);
    for (uint32_t q = 0; q < sg.n; q++) if (flt_check_boundary(t, sg.g[q], a1.chr, a1.pos)) return true;
    return false;
}

template <bool EMIT>
__global__ void __launch_bounds__(256) splice_kernel(const SpliceArgs a)
{
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = lane_id();
    FltSeg *segs = (FltSeg *)(a.scratch + (size_t)warp * a.scratch_per_warp);
    SpliceGenes *genes = (SpliceGenes *)(segs + a.seg_cap);
    const FltTables &t = a.t;
    for (;;) {
        const uint32_t i = fetch_work(a.work);
        if (i >= a.n) break;
        const int which = a.pair_needs_host[i] ? 0 : a.ev[i].unaligned;
        if (!which) continue;
        if (EMIT && (a.overflow[i] || a.kind[i] == FLT_SPLICE_NONE)) continue;
        const int e = which - 1;
        const uint32_t read_len = a.offsets[e][i + 1] - a.offsets[e][i];
        const unsigned long long lo = a.seg[e][2 * (size_t)i], mid = a.seg[e][2 * (size_t)i + 1], hi = a.seg[e][2 * (size_t)i + 2];
        // the partial alignments, in the order of the reference's vector: forward map ascending, then RC map ascending
        uint32_t n = 0;
        bool over = false;
        #pragma unroll 1
        for (unsigned long long q0 = lo; q0 < hi && !over; q0 += 32) {
            const unsigned long long q = q0 + lane;
            bool first = false;
            FltSeg s;
            if (q < hi) {
                const bool rc = q >= mid;
                const unsigned long long begin = rc ? mid : lo, end = rc ? hi : mid;
                const uint32_t loc = a.ch_loc[e][q];
                first = q == begin || a.ch_loc[e][q - 1] != loc;
                if (first) {
                    unsigned long long last = q;
                    while (last + 1 < end && a.ch_loc[e][last + 1] == loc) last++;
                    const int p = flt_piece_at(t.piece_begin, (int)t.n_pieces, loc);
                    if (p < 0) first = false;
                    else {
                        const uint32_t o0 = a.ch_off[e][q], o1 = a.ch_off[e][last], length = (o1 - o0) + a.seed_len;
                        const int32_t pos0 = (int32_t)(loc - t.piece_begin[p] + 1);
                        const uint32_t start = rc ? (uint32_t)pos0 + read_len - (o1 + a.seed_len) : (uint32_t)pos0 + o0;
                        s.chr = p; s.pos = start; s.pos_end = start + length - 1; s.score = length;
                    }
                }
            }
            const uint32_t b = __ballot_sync(FULL_MASK, first);
            const uint32_t at = n + __popc(b & ((1u << lane) - 1));
            over = n + __popc(b) > a.seg_cap;
            if (first && !over) segs[at] = s;
            n += __popc(b);
        }
        if (over) {
            if (!EMIT && lane == 0) { a.overflow[i] = 1; a.counts[i] = 0; a.kind[i] = FLT_SPLICE_NONE; }
            continue;
        }
        __syncwarp();
        // the genes that cover each alignment (GTFReader::IntervalGenes), all lanes over the gene table
        #pragma unroll 1
        for (uint32_t k = 0; k < n; k++) {
            const FltSeg s = segs[k];
            uint32_t cnt = 0;
            #pragma unroll 1
            for (uint32_t g0 = 0; g0 < t.n_genes; g0 += 32) {
                const uint32_t g = g0 + lane;
                const bool hit = g < t.n_genes && flt_gene_found(t, g, s.chr, s.pos, s.pos_end);
                const uint32_t b = __ballot_sync(FULL_MASK, hit);
                if (hit) { const uint32_t at = cnt + __popc(b & ((1u << lane) - 1)); if (at < SPLICE_GENES) genes[k].g[at] = (int32_t)g; }
                cnt += __popc(b);
            }
            if (lane == 0) genes[k].n = cnt;
        }
        __syncwarp();
        // the candidate loop (AlignmentFilter.cpp:809-876)
        const int want = EMIT ? (int)a.kind[i] : 0;
        unsigned long long c_intra = 0, c_inter = 0, w = EMIT ? a.counts[i] : 0;
        bool gene = false;
        #pragma unroll 1
        for (uint32_t k = 0; k + 1 < n && !gene; k++) {
            const FltSeg s0 = segs[k];
            const SpliceGenes sg = genes[k];
            #pragma unroll 1
            for (uint32_t j0 = k + 1; j0 < n; j0 += 32) {
                const uint32_t j = j0 + lane;
                int c = FLT_SPLICE_NONE;
                FltSeg s1;
                if (j < n) {
                    s1 = segs[j];
                    c = flt_splice_class(s0, s1, read_len, a.seed_len);
                    if (!EMIT && c == FLT_SPLICE_INTRACHR && splice_in_gene(t, s0, sg, s1)) c = FLT_SPLICE_GENE;
                }
                if (!EMIT) {
                    if (__any_sync(FULL_MASK, c == FLT_SPLICE_GENE)) { gene = true; break; }
                    c_intra += __popc(__ballot_sync(FULL_MASK, c == FLT_SPLICE_INTRACHR));
                    c_inter += __popc(__ballot_sync(FULL_MASK, c == FLT_SPLICE_INTERCHR));
                } else {
                    const uint32_t b = __ballot_sync(FULL_MASK, c == want);
                    if (c == want) {
                        FltSplice &o = a.out[w + __popc(b & ((1u << lane) - 1))];
                        o.pair = i; o.kind = c;
                        o.chr[0] = s0.chr; o.pos[0] = s0.pos; o.pos_end[0] = s0.pos_end;
                        o.chr[1] = s1.chr; o.pos[1] = s1.pos; o.pos_end[1] = s1.pos_end;
                    }
                    w += __popc(b);
                }
            }
        }
        if (!EMIT && lane == 0) {
            const int kind = gene ? FLT_SPLICE_NONE : c_intra ? FLT_SPLICE_INTRACHR : c_inter ? FLT_SPLICE_INTERCHR : FLT_SPLICE_NONE;
            a.kind[i] = (uint8_t)kind;
            a.overflow[i] = 0;
            a.counts[i] = kind == FLT_SPLICE_INTRACHR ? c_intra : kind == FLT_SPLICE_INTERCHR ? c_inter : 0;
        }
        __syncwarp();
    }
}
