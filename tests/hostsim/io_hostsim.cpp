// io_hostsim.cpp -- TEST INFRASTRUCTURE ONLY.  Runs the per-record logic of the FASTQ / SAM batch kernels
// (snap_rnaseq_b200/csrc/iofmt.h, the header the device code compiles) on the host, so that `pytest -m "not gpu"` can
// check it against the compiled reference where there is no GPU.  Nothing in the product loads this; the warp-parallel
// parts of the kernels (newline scan, QNAME/NUL scans, the CIGAR walk) are replaced by serial loops here, and the CIGAR
// strings are handed in by the test (from the reference's own LandauVishkinWithCigar).
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../../snap_rnaseq_b200/csrc/iofmt.h"

struct HostsimIndex {
    const uint32_t *piece_begin;
    uint32_t n_pieces;
    const char *names_blob;
    const uint32_t *names_off;
    const char *cigars[2];       // [n][stride] per end, for the clipped read at its (masked) location and direction
    const int32_t *edit_distance[2];  // computeEditDistance's return value; -3: off the genome ("*", NM -1)
    uint32_t cigar_stride;
};

extern "C" int hostsim_fastq_parse(const uint8_t *text_in, uint64_t n_bytes, int clipping, uint32_t max_reads, uint32_t *n_reads,
                                   uint64_t *bytes_consumed, uint32_t *offsets, uint8_t *bases, uint8_t *quals, uint16_t *front_clip,
                                   uint16_t *clipped_len, uint32_t *id_offsets, uint8_t *ids)
{
    std::vector<uint8_t> text(text_in, text_in + n_bytes);
    text.resize(n_bytes + 65536 + 64, 0);
    std::vector<uint32_t> nl;
    for (uint64_t i = 0; i < n_bytes; i++) if (text[i] == '\n') nl.push_back((uint32_t)i);
    const uint32_t n = (uint32_t)(nl.size() / 4);
    *n_reads = 0;
    *bytes_consumed = 0;
    offsets[0] = id_offsets[0] = 0;
    if (n > max_reads) return -4;
    for (uint32_t r = 0; r < n; r++) {
        const FqRecord rec = fq_record(text.data(), n_bytes, nl.data(), r, clipping);
        if (rec.error) return rec.error == FQ_TOO_LONG ? -4 : -1;
        const uint32_t o = offsets[r], io = id_offsets[r];
        for (uint32_t i = 0; i < rec.data_len; i++) {
            bases[o + i] = fq_upper(text[rec.data_start + i]);
            quals[o + i] = text[rec.qual_start + i];
        }
        memcpy(ids + io, text.data() + rec.id_start, rec.id_len);
        offsets[r + 1] = o + rec.data_len;
        id_offsets[r + 1] = io + rec.id_len;
        front_clip[r] = rec.front_clip;
        clipped_len[r] = rec.clipped_len;
        *bytes_consumed = rec.end;
    }
    *n_reads = n;
    return 0;
}

extern "C" int hostsim_fastq_record_start(const uint8_t *text, uint64_t n_bytes, uint64_t *offset)
{
    *offset = n_bytes ? fq_record_start(text, n_bytes) : 0;
    return 0;
}

static SamReadsDev view(const snapb200_sam_reads *r)
{
    SamReadsDev d = {r->offsets, r->bases, r->quals, r->front_clip, r->clipped_len, r->id_offsets, r->ids};
    return d;
}

extern "C" int hostsim_sam_batch(const HostsimIndex *ix, const snapb200_sam_reads *reads0, const snapb200_sam_reads *reads1,
                                 const snapb200_sam_alignment *aln0, const snapb200_sam_alignment *aln1, int use_m, const char *read_group,
                                 char *out, uint64_t out_capacity, uint64_t *line_offsets)
{
    const bool bam = (use_m & SNAPB200_SAM_BAM_RECORDS) != 0;
    SamInputs in;
    in.paired = reads1 != NULL;
    in.rd[0] = view(reads0);
    in.aln[0] = aln0;
    if (in.paired) { in.rd[1] = view(reads1); in.aln[1] = aln1; }
    const SamNames names = {ix->names_blob, ix->names_off};
    const uint32_t rg_len = read_group ? (uint32_t)strlen(read_group) : 0;
    const uint32_t n_lines = reads0->n * (in.paired ? 2 : 1);
    uint64_t pos = 0;
    line_offsets[0] = 0;
    for (uint32_t line = 0; line < n_lines; line++) {
        const SamWho w = sam_who(in, line);
        if (w.skip) { line_offsets[line + 1] = pos; continue; }
        const SamReadsDev &rd = in.rd[w.e];
        const uint32_t off = rd.offsets[w.i];
        const uint8_t *bases = rd.bases + off, *quals = rd.quals + off, *id = rd.ids + rd.id_offsets[w.i];
        const SamFields f = sam_fields(ix->piece_begin, (int)ix->n_pieces, w.me, w.has_mate, w.first_in_pair, w.mate);
        SamLine ln;
        uint32_t qn = rd.id_offsets[w.i + 1] - rd.id_offsets[w.i];
        if (in.paired) {
            const SamReadsDev &r0 = in.rd[0], &r1 = in.rd[1];
            if (sam_pair_trims_ids(r0.ids + r0.id_offsets[w.i], r0.id_offsets[w.i + 1] - r0.id_offsets[w.i], r1.ids + r1.id_offsets[w.i],
                                   r1.id_offsets[w.i + 1] - r1.id_offsets[w.i]))
                qn -= 2;
        }
        ln.qname_len = qn;
        ln.spliced = ln.n_ops = ln.ref_len = 0;
        for (uint32_t i = 0; i < qn && !bam; i++) if (id[i] == ' ') { ln.qname_len = i; break; }
        ln.seq_len = ln.qual_len = w.me.full_len;
        for (uint32_t i = 0; i < w.me.full_len; i++) {
            const uint8_t b = f.direction == 1 ? bases[w.me.full_len - 1 - i] : bases[i];
            const bool z = f.direction == 1 ? !(b == 'A' || b == 'C' || b == 'G' || b == 'T' || b == 'N' || b == 'n') : b == 0;
            if (z) { ln.seq_len = i; break; }
        }
        for (uint32_t i = 0; i < w.me.full_len; i++) if ((f.direction == 1 ? quals[w.me.full_len - 1 - i] : quals[i]) == 0) { ln.qual_len = i; break; }
        ln.edit_distance = -1;
        ln.cigar_len = 0;
        const char *cig = ix->cigars[w.e] + (size_t)w.i * ix->cigar_stride;
        if (f.mapped) {
            const int e = ix->edit_distance[w.e][w.i];
            if (e != -3) {
                ln.edit_distance = e;
                if (e >= 0) ln.cigar_len = sam_strlen(cig, ix->cigar_stride);
            }
        }
        if (bam) {
            ln.seq_len = ln.qual_len = w.me.full_len;
            bam_count_ops(f, &ln, cig);
            const uint32_t len = bam_record_len(ln, w.me.full_len, rg_len);
            if (out) {
                if (pos + len > out_capacity) return -1;
                uint8_t *rec = (uint8_t *)out + pos;
                uint8_t *seq = bam_put_head(rec, len, id, f, ln, w.me.full_len, cig);
                if (seq != rec + 36 + ln.qname_len + 1 + 4 * ln.n_ops) return -101;
                bam_put_seq_qual(seq, bases, quals, w.me.full_len, f.direction, 0, 1);
                uint8_t *end = bam_put_aux(seq + (w.me.full_len + 1) / 2 + w.me.full_len, ln, read_group, rg_len);
                if ((uint32_t)(end - rec) != len) return -100;
            }
            pos += len;
            line_offsets[line + 1] = pos;
            continue;
        }
        const uint32_t len = sam_line_len(f, ln, names, rg_len);
        if (out) {
            if (pos + len > out_capacity) return -1;
            char *dst = out + pos;
            char *seq = sam_put_prefix(dst, id, f, ln, names, cig);
            sam_put_seq_qual(seq, bases, quals, w.me.full_len, f.direction, ln, 0, 1);
            char *end = sam_put_suffix(seq + ln.seq_len + 1 + ln.qual_len, ln, read_group, rg_len);
            if ((uint32_t)(end - dst) != len) return -100;  // the measuring pass and the writing pass must agree
        }
        pos += len;
        line_offsets[line + 1] = pos;
    }
    return 0;
}

// ---- row f3 groundwork: the alignment lists of one pair, as AlignmentFilter builds them ------------------------------------------
#include "../../snap_rnaseq_b200/csrc/filterfmt.h"

extern "C" int hostsim_filter_alignments(const FltTables *t, uint32_t len0, uint32_t len1, uint32_t max_dist, int n0, const uint32_t *l0,
                                         const uint8_t *rc0, const int32_t *sc0, int n1, const uint32_t *l1, const uint8_t *rc1, const int32_t *sc1,
                                         const snapb200_paired_result *g, uint32_t cap, uint32_t *counts, uint32_t *records)
{
    std::vector<FltAln> lists[2];
    lists[0].resize(cap + 1);
    lists[1].resize(cap + 1);
    uint32_t n[2] = {0, 0};
    FltAln a;
    for (int k = 0; k < n0; k++) if (flt_make_alignment(*t, l0[k], rc0[k] ? 1 : 0, sc0[k], 0, true, len0, max_dist, &a)) { if (n[0] >= cap) return -2; n[0] = flt_insert(*t, lists[0].data(), n[0], a); }
    for (int k = 0; k < n1; k++) if (flt_make_alignment(*t, l1[k], rc1[k] ? 1 : 0, sc1[k], 0, true, len1, max_dist, &a)) { if (n[1] >= cap) return -2; n[1] = flt_insert(*t, lists[1].data(), n[1], a); }
    const uint32_t lens[2] = {len0, len1};
    for (int e = 0; e < 2; e++)
        if (flt_make_alignment(*t, g->location[e], g->direction[e], g->score[e], g->mapq[e], false, lens[e], max_dist, &a)) { if (n[e] >= cap) return -2; n[e] = flt_insert(*t, lists[e].data(), n[e], a); }
    for (int e = 0; e < 2; e++) {
        counts[e] = n[e];
        for (uint32_t c = 0; c < n[e]; c++) {
            const FltAln &x = lists[e][c];
            uint32_t *r = records + ((size_t)e * cap + c) * 7;
            r[0] = x.location; r[1] = x.pos; r[2] = x.pos_end; r[3] = x.pos_original; r[4] = (uint32_t)x.score; r[5] = x.direction; r[6] = x.is_transcriptome;
        }
    }
    return 0;
}

// AlignmentFilter::Filter for one pair on the host: lists as above, every combination classified in the reference's loop order,
// flt_sort_pairs (the mirror of libstdc++'s std::sort, checked against std::sort below) for ProcessPairs.
#include <algorithm>
extern "C" int hostsim_filter_pair(const FltTables *t, uint32_t len0, uint32_t len1, uint32_t max_dist, uint32_t max_spacing, uint32_t conf_diff,
                                   int force_spacing, int n0, const uint32_t *l0, const uint8_t *rc0, const int32_t *sc0, int n1, const uint32_t *l1,
                                   const uint8_t *rc1, const int32_t *sc1, const snapb200_paired_result *g, const uint64_t *seg0, const uint32_t *clocs0,
                                   const uint16_t *coffs0, const uint64_t *seg1, const uint32_t *clocs1, const uint16_t *coffs1, uint32_t pair_index,
                                   FltResult *out, FltEvent *ev)
{
    const uint32_t cap = 2048;
    std::vector<FltAln> lists[2];
    lists[0].resize(cap + 1);
    lists[1].resize(cap + 1);
    uint32_t n[2] = {0, 0};
    FltAln a;
    for (int k = 0; k < n0; k++) if (flt_make_alignment(*t, l0[k], rc0[k] ? 1 : 0, sc0[k], 0, true, len0, max_dist, &a)) n[0] = flt_insert(*t, lists[0].data(), n[0], a);
    for (int k = 0; k < n1; k++) if (flt_make_alignment(*t, l1[k], rc1[k] ? 1 : 0, sc1[k], 0, true, len1, max_dist, &a)) n[1] = flt_insert(*t, lists[1].data(), n[1], a);
    const uint32_t lens[2] = {len0, len1};
    for (int e = 0; e < 2; e++)
        if (flt_make_alignment(*t, g->location[e], g->direction[e], g->score[e], g->mapq[e], false, lens[e], max_dist, &a)) n[e] = flt_insert(*t, lists[e].data(), n[e], a);
    std::vector<FltPair> cls[4];
    for (uint32_t j = 0; j < n[1]; j++)        // the reference's outer loop runs over its `mate0` map = read 1's alignments
        for (uint32_t i = 0; i < n[0]; i++)
            cls[flt_classify(*t, lists[1][j], lists[0][i])].push_back(flt_make_pair(lists[0][i], lists[1][j], i, j));
    FltResult r;
    memset(&r, 0, sizeof(r));
    for (int e = 0; e < 2; e++) {  // what the run loop hands in (PairedAligner.cpp:575-578, 620)
        r.location[e] = g->location[e]; r.score[e] = g->score[e]; r.mapq[e] = g->mapq[e]; r.status[e] = g->status[e]; r.direction[e] = g->direction[e];
    }
    uint32_t genome_mapq = 70;
    memset(ev, 0, sizeof(*ev));
    ev->transcript[0] = ev->transcript[1] = -1;
    if (n[0] == 0 && n[1] != 0) ev->unaligned = 1;  // the reference's mate1 map (read 0's alignments) is empty: UnalignedRead(read0)
    if (n[1] == 0 && n[0] != 0) ev->unaligned = 2;
    auto event = [&](int kind, const FltPair &p) {
        ev->kind = kind;
        const FltAln *al[2] = {&lists[0][p.a1], &lists[1][p.a2]};
        for (int e = 0; e < 2; e++) {
            ev->transcript[e] = al[e]->transcript; ev->chr[e] = al[e]->chr;
            ev->pos_original[e] = al[e]->pos_original; ev->pos[e] = al[e]->pos; ev->pos_end[e] = al[e]->pos_end;
        }
    };
    auto partial = [&]() {
        std::vector<uint32_t> p0(1 << 16), p1(1 << 16);
        uint32_t c0 = 0, c1 = 0;
        const uint64_t s = 2ull * pair_index;  // snapb200_characterize_batch layout: segment 2 * read + direction
        flt_partial_locations(clocs0, coffs0, seg0[s], seg0[s + 1], false, len0, p0.data(), &c0);
        flt_partial_locations(clocs0, coffs0, seg0[s + 1], seg0[s + 2], true, len0, p0.data(), &c0);
        flt_partial_locations(clocs1, coffs1, seg1[s], seg1[s + 1], false, len1, p1.data(), &c1);
        flt_partial_locations(clocs1, coffs1, seg1[s + 1], seg1[s + 2], true, len1, p1.data(), &c1);
        if (flt_partial_match(*t, p0.data(), c0, p1.data(), c1, max_spacing)) { r.status[0] = r.status[1] = 2; r.mapq[0] = r.mapq[1] = 1; }
    };
    auto process = [&](std::vector<FltPair> &v) {
        if (v.size() > 1) flt_sort_pairs(v.data(), (long)v.size());
        flt_process_pairs(*t, lists[0].data(), lists[1].data(), v.data(), (uint32_t)v.size(), conf_diff, &genome_mapq, &r);
    };
    if (!cls[FLT_INTRAGENE].empty()) {
        process(cls[FLT_INTRAGENE]);
        r.aligned_as_pair = 1;  // AlignmentFilter.cpp:543-548; every other exit of Filter leaves alignedAsPair false
        if (r.status[0] == 1) event(FLT_EV_INCREMENT, cls[FLT_INTRAGENE][0]);
    } else if (!cls[FLT_INTRACHR].empty()) {
        process(cls[FLT_INTRACHR]);
        if (r.status[0] == 1) flt_check_no_rc(lists[0].data(), lists[1].data(), cls[FLT_NO_RC].data(), (uint32_t)cls[FLT_NO_RC].size(), &r);
        if (!((uint32_t)cls[FLT_INTRACHR][0].distance <= max_spacing)) {
            if (r.status[0] == 1) partial();
            if (r.status[0] == 1) event(FLT_EV_INTRACHR, cls[FLT_INTRACHR][0]);
        }
    } else if (!cls[FLT_INTERCHR].empty()) {
        process(cls[FLT_INTERCHR]);
        if (r.status[0] == 1) flt_check_no_rc(lists[0].data(), lists[1].data(), cls[FLT_NO_RC].data(), (uint32_t)cls[FLT_NO_RC].size(), &r);
        if (r.status[0] == 1) partial();
        if (r.status[0] == 1) event(FLT_EV_INTERCHR, cls[FLT_INTERCHR][0]);
    } else if (!cls[FLT_NO_RC].empty()) {
        process(cls[FLT_NO_RC]);
        if (r.status[0] == 1) partial();
        if (r.status[0] == 1) {
            const FltPair &p0 = cls[FLT_NO_RC][0];
            event(lists[0][p0.a1].chr == lists[1][p0.a2].chr ? FLT_EV_INTRACHR : FLT_EV_INTERCHR, p0);
        }
    } else {
        memset(&r, 0, sizeof(r));  // NotFound, location 0 (not InvalidGenomeLocation), FORWARD (AlignmentFilter.cpp:717-735)
    }
    const bool one0 = r.status[0] == 1, one1 = r.status[1] == 1;  // isOneLocation: SingleHit (CertainHit does not occur here)
    if (force_spacing && one0 != one1) { r.status[0] = r.status[1] = 0; r.location[0] = r.location[1] = 0xffffffffu; }
    if (r.score[0] + r.score[1] >= 5) {  // "cheese", PairedAligner.cpp:653-663
        if (r.mapq[0] < 50) r.mapq[0] /= 2;
        if (r.mapq[1] < 50) r.mapq[1] /= 2;
    }
    for (int e = 0; e < 2; e++) if (!r.is_transcriptome[e]) r.tlocation[e] = 0;
    *out = r;
    return 0;
}


// flt_sort_pairs against std::sort on the same sequence of scores: the permutations must be identical, ties included
extern "C" int hostsim_sort_check(const uint32_t *scores, uint32_t n, uint32_t *mine, uint32_t *theirs)
{
    std::vector<FltPair> a(n & 0x7fffffffu), b(n & 0x7fffffffu);
    for (uint32_t i = 0; i < (n & 0x7fffffffu); i++) { a[i].a1 = i; a[i].a2 = 0; a[i].distance = 0; a[i].score = scores[i]; b[i] = a[i]; }
    if (n & 0x80000000u) {  // high bit: the heapsort fallback of introsort alone, against std::partial_sort(first, last, last)
        n &= 0x7fffffffu;
        a.resize(n); b.resize(n);
        flt_heap_sort(a.data(), (long)n);
        std::partial_sort(b.begin(), b.end(), b.end());
    } else {
        flt_sort_pairs(a.data(), (long)n);
        std::sort(b.begin(), b.end());
    }
    for (uint32_t i = 0; i < n; i++) { mine[i] = a[i].a1; theirs[i] = b[i].a1; }
    return 0;
}

// The annotation loader (gtf_tables.h) written out in the format of oracle/ref_driver.cpp::ref_gtf_export, for a byte comparison
#include "../../snap_rnaseq_b200/csrc/gtf_tables.h"
#include <stdio.h>
extern "C" int hostsim_gtf_export(const char *gtf_path, const char *out_path)
{
    GtfTables t;
    if (!gtf_load_tables(gtf_path, &t)) return -1;
    FILE *f = fopen(out_path, "w");
    if (!f) return -2;
    for (size_t i = 0; i < t.transcripts.size(); i++) {
        const GtfTranscriptRow &r = t.transcripts[i];
        fprintf(f, "T\t%s\t%s\t%s\t%u\t%u\t%u", r.id.c_str(), r.chr.c_str(), r.gene_id.c_str(), r.start, r.end, (unsigned)r.features.size());
        for (size_t k = 0; k < r.features.size(); k++) fprintf(f, "\t%u\t%u\t%u", r.features[k].type, r.features[k].start, r.features[k].end);
        fprintf(f, "\n");
    }
    for (size_t i = 0; i < t.genes.size(); i++) fprintf(f, "G\t%s\t%s\t%u\t%u\n", t.genes[i].id.c_str(), t.genes[i].chr.c_str(), t.genes[i].start, t.genes[i].end);
    fclose(f);
    return 0;
}


// The same pair through flt_filter_pair: the allocation-free form a kernel will run (fixed scratch, no STL)
extern "C" int hostsim_filter_pair_flat(const FltTables *t, uint32_t len0, uint32_t len1, uint32_t max_dist, uint32_t max_spacing, uint32_t conf_diff,
                                        int force_spacing, int n0, const uint32_t *l0, const uint8_t *rc0, const int32_t *sc0, int n1, const uint32_t *l1,
                                        const uint8_t *rc1, const int32_t *sc1, const snapb200_paired_result *g, const uint64_t *seg0, const uint32_t *clocs0,
                                        const uint16_t *coffs0, const uint64_t *seg1, const uint32_t *clocs1, const uint16_t *coffs1, uint32_t pair_index,
                                        uint32_t list_cap, uint32_t pair_cap, uint32_t ploc_cap, FltResult *out, FltEvent *ev)
{
    std::vector<FltAln> la(list_cap + 1), lb(list_cap + 1);
    std::vector<FltPair> pairs(pair_cap + 1);
    std::vector<uint32_t> pa(ploc_cap + 1), pb(ploc_cap + 1);
    FltScratch sc;
    sc.list[0] = la.data(); sc.list[1] = lb.data(); sc.list_cap = list_cap;
    sc.pairs = pairs.data(); sc.pair_cap = pair_cap;
    sc.ploc[0] = pa.data(); sc.ploc[1] = pb.data(); sc.ploc_cap = ploc_cap;
    FltParams prm = {max_dist, max_spacing, conf_diff, force_spacing};
    FltPairInput in;
    in.len[0] = len0; in.len[1] = len1;
    in.n_hits[0] = n0; in.n_hits[1] = n1;
    in.hit_loc[0] = l0; in.hit_loc[1] = l1; in.hit_rc[0] = rc0; in.hit_rc[1] = rc1; in.hit_score[0] = sc0; in.hit_score[1] = sc1;
    for (int e = 0; e < 2; e++) {
        in.g_location[e] = g->location[e]; in.g_score[e] = g->score[e]; in.g_mapq[e] = g->mapq[e]; in.g_status[e] = g->status[e]; in.g_direction[e] = g->direction[e];
    }
    const uint64_t s = 2ull * pair_index;
    in.ch_loc[0] = clocs0; in.ch_off[0] = coffs0; in.ch_loc[1] = clocs1; in.ch_off[1] = coffs1;
    for (int k = 0; k < 3; k++) { in.ch_range[0][k] = seg0[s + k]; in.ch_range[1][k] = seg1[s + k]; }
    return flt_filter_pair(*t, prm, in, sc, out, ev);
}


// AlignmentFilter::UnalignedRead as records (flt_unaligned_segments + flt_unaligned_splices, the serial specification): the splice
// records of read `which` (1: read 0, 2: read 1) of one pair.  out == NULL: count.  Returns the count, -1 if there are more than
// seg_cap partial alignments.
extern "C" long long hostsim_unaligned_splices(const FltTables *t, uint32_t read_len, uint32_t seed_len, const uint64_t *seg, const uint32_t *clocs,
                                               const uint16_t *coffs, uint32_t pair_index, uint32_t seg_cap, FltSplice *out)
{
    std::vector<FltSeg> segs(seg_cap + 1);
    const uint64_t s = 2ull * pair_index;
    const int n = flt_unaligned_segments(*t, clocs, coffs, seg[s], seg[s + 1], seg[s + 2], read_len, seed_len, segs.data(), seg_cap);
    if (n < 0) return -1;
    int kind = 0;
    return (long long)flt_unaligned_splices(*t, segs.data(), (uint32_t)n, read_len, seed_len, pair_index, &kind, out);
}

// ---- the CIGAR of a transcriptome alignment (iofmt.h: sam_splice_cigar) -----------------------------------------------------------
extern "C" int hostsim_splice_cigar(const FltTables *t, int tr, uint32_t pos, const char *lv, uint32_t lv_len, uint32_t clip_before, uint32_t clip_after,
                                    char *out, uint32_t cap, uint32_t *n_calls)
{
    return sam_splice_cigar(*t, tr, pos, lv, lv_len, clip_before, clip_after, out, cap, n_calls);
}

// ---- BGZF blocks (bgzf.h): the serial specification, and the pieces the kernel combines --------------------------------------------
#include "../../snap_rnaseq_b200/csrc/bgzf.h"
#include <vector>

extern "C" long long hostsim_bgzf_compress(const uint8_t *in, uint64_t n, uint32_t chunk, uint8_t *out, uint64_t cap)
{
    uint32_t table[256];
    for (uint32_t i = 0; i < 256; i++) table[i] = bgzf_crc_table_entry(i);
    std::vector<uint32_t> scratch(257 + 257 + 513 + 513);
    uint64_t pos = 0;
    for (uint64_t lo = 0; lo < n || (n == 0 && lo == 0); lo += chunk) {
        const uint32_t m = (uint32_t)std::min<uint64_t>(chunk, n - lo);
        if (pos + BGZF_HEADER + m + 5 + BGZF_FOOTER > cap) return -1;
        memset(out + pos, 0, BGZF_HEADER + m + 5 + BGZF_FOOTER);
        pos += bgzf_block_serial(table, in + lo, m, out + pos, scratch.data());
        if (n == 0) break;
    }
    return (long long)pos;
}

// the CRC of a buffer computed the way 32 lanes do: slices with their own registers, combined through the zero-shift matrices
extern "C" uint32_t hostsim_bgzf_crc_sliced(const uint8_t *in, uint32_t n, uint32_t slices)
{
    uint32_t table[256], shift[17 * 32];
    for (uint32_t i = 0; i < 256; i++) table[i] = bgzf_crc_table_entry(i);
    bgzf_crc_shift_build(table, shift);
    const uint32_t per = (n + slices - 1) / slices;
    uint32_t total = 0;
    for (uint32_t k = 0; k < slices; k++) {
        const uint32_t lo = std::min(n, k * per), hi = std::min(n, lo + per);
        const uint32_t r = bgzf_crc_raw(table, k == 0 ? 0xffffffffu : 0u, in + lo, hi - lo);
        total ^= bgzf_crc_zeros(shift, r, n - hi);
    }
    return total ^ 0xffffffffu;
}
