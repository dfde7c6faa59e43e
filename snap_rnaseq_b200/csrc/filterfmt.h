// filterfmt.h -- the per-element rules of the device AlignmentFilter as plain host/device functions
// over flat tables, written the way iofmt.h is: filter_warp.cuh (the warp-per-pair kernel behind snapb200_filter_paired_batch and
// snapb200_rna_batch_*) calls them on the device, and tests/hostsim runs the same header on the host against the compiled reference.
//   flt_make_alignment  AlignmentFilter::AddAlignment      (SNAPLib/AlignmentFilter.cpp:140-214)
//   flt_genomic_position GTFTranscript::GenomicPosition    (SNAPLib/GTFReader.cpp:1075-1107)
//   flt_key_compare     the order of std::map<std::string, Alignment> keyed rname + '_' + ToString(pos) (AlignmentFilter.cpp:29)
//   flt_insert          AlignmentFilter::HashAlignment     (SNAPLib/AlignmentFilter.cpp:113-138)
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define FLT_HD __host__ __device__ __forceinline__
#else
#define FLT_HD static inline
#endif

#define FLT_INVALID_LOC 0xffffffffu
#define FLT_EXON 1  // enum {UNASSIGNED, EXON, INTRON}, SNAPLib/GTFReader.h:49

struct FltTables {
    // genome and transcriptome pieces (Genome::Piece::beginningOffset); chromosome names for the map keys
    const uint32_t *piece_begin; uint32_t n_pieces;
    const char *chr_names; const uint32_t *chr_name_off;      // [n_pieces + 1]
    const uint32_t *tpiece_begin; uint32_t n_tpieces;
    const int32_t *tpiece_transcript;                          // transcript of each transcriptome piece (GTFReader::GetTranscript(piece name))
    // transcripts: chromosome (as a genome piece index), gene, end, and their features in GTFTranscript::exons order
    const int32_t *t_chr, *t_gene; const uint32_t *t_end; const uint32_t *t_feat_first;  // [n_transcripts + 1]
    const uint32_t *f_type, *f_start, *f_end;
    // genes: chromosome (genome piece index) and extent, in gene-id order (the reference's gene map)
    const int32_t *g_chr; const uint32_t *g_start, *g_end;
    uint32_t n_genes;
    int32_t gene_tree_min_stop;  // see flt_gene_found
};

struct FltAln {  // class Alignment, SNAPLib/AlignmentFilter.h:41-69, with names as indices
    uint32_t location, pos, pos_end, pos_original;
    int32_t score, mapq;
    int32_t chr;         // genome piece index of rname
    int32_t transcript;  // -1 for a genome alignment
    int32_t gene;
    uint8_t direction, is_transcriptome;
};

FLT_HD int flt_piece_at(const uint32_t *piece_begin, int n_pieces, uint32_t location)
{  // Genome::getPieceAtLocation, SNAPLib/Genome.cpp:357-374
    int low = 0, high = n_pieces - 1;
    while (low <= high) {
        const int mid = (low + high) / 2;
        if (piece_begin[mid] <= location && (mid == n_pieces - 1 || piece_begin[mid + 1] > location)) return mid;
        if (piece_begin[mid] <= location) low = mid + 1;
        else high = mid - 1;
    }
    return -1;
}

FLT_HD uint32_t flt_genomic_position(const FltTables &t, int tr, uint32_t transcript_pos, uint32_t span)
{
    for (uint32_t k = t.t_feat_first[tr]; k < t.t_feat_first[tr + 1]; k++) {
        if (t.f_type[k] != FLT_EXON) continue;
        const uint32_t len = t.f_end[k] - t.f_start[k] + 1;
        if (transcript_pos > len) {
            transcript_pos -= len;
        } else {
            const uint32_t genome_pos = t.f_start[k] + transcript_pos - 1;
            if (genome_pos + span > t.t_end[tr]) return 0;  // the read runs past the transcript (consecutive pieces in the index)
            return genome_pos;
        }
    }
    return 0;
}

// Returns false when AddAlignment adds nothing (score gate, no location, position 0).  own_len = the data length of the read the
// alignment belongs to (the run loop's isMate0 flag is inverted, which makes the reference use exactly that).
FLT_HD bool flt_make_alignment(const FltTables &t, uint32_t location, int direction, int score, int mapq, bool is_transcriptome, uint32_t own_len,
                               uint32_t max_dist, FltAln *out)
{
    if ((uint32_t)score > max_dist) return false;  // int against unsigned, as in the reference: a negative score is dropped too
    if (location == FLT_INVALID_LOC) return false;
    FltAln a;
    a.location = location; a.direction = (uint8_t)direction; a.score = score; a.mapq = mapq; a.is_transcriptome = is_transcriptome;
    a.transcript = -1; a.gene = -1;
    if (!is_transcriptome) {
        const int p = flt_piece_at(t.piece_begin, (int)t.n_pieces, location);
        if (p < 0) return false;
        a.chr = p;
        a.pos_original = location - t.piece_begin[p] + 1;
        a.pos = a.pos_original;
        a.pos_end = a.pos + own_len - 1;
    } else {
        const int p = flt_piece_at(t.tpiece_begin, (int)t.n_tpieces, location);
        if (p < 0) return false;
        const int tr = t.tpiece_transcript[p];
        a.transcript = tr;
        a.gene = t.t_gene[tr];
        a.chr = t.t_chr[tr];
        a.pos_original = location - t.tpiece_begin[p] + 1;
        a.pos_end = flt_genomic_position(t, tr, a.pos_original + own_len - 1, 0);
        a.pos = flt_genomic_position(t, tr, a.pos_original, own_len);
    }
    if (a.pos == 0) return false;
    *out = a;
    return true;
}

// strcmp of the keys rname + '_' + decimal(pos) without building them: names first, byte by byte, with '_' and the digits standing
// in once a name ends ("chr1_5" against "chr10_5" is decided by '_' against '0').
FLT_HD int flt_key_char(const FltTables &t, int chr, uint32_t pos, uint32_t i, const char *digits, int nd)
{
    const uint32_t nl = t.chr_name_off[chr + 1] - t.chr_name_off[chr];
    if (i < nl) return (unsigned char)t.chr_names[t.chr_name_off[chr] + i];
    if (i == nl) return '_';
    const uint32_t k = i - nl - 1;
    return k < (uint32_t)nd ? digits[k] : 0;
}

FLT_HD int flt_decimal(uint32_t v, char *buf)
{
    char tmp[10];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    for (int i = 0; i < n; i++) buf[i] = tmp[n - 1 - i];
    return n;
}

FLT_HD int flt_key_compare(const FltTables &t, int chr_a, uint32_t pos_a, int chr_b, uint32_t pos_b)
{
    if (chr_a == chr_b && pos_a == pos_b) return 0;
    char da[10], db[10];
    const int na = flt_decimal(pos_a, da), nb = flt_decimal(pos_b, db);
    for (uint32_t i = 0;; i++) {
        const int ca = flt_key_char(t, chr_a, pos_a, i, da, na), cb = flt_key_char(t, chr_b, pos_b, i, db, nb);
        if (ca != cb) return ca < cb ? -1 : 1;
        if (ca == 0) return 0;
    }
}

// HashAlignment into a list kept in key order; returns the new count.  The list must have room for one more.
FLT_HD uint32_t flt_insert(const FltTables &t, FltAln *list, uint32_t n, const FltAln &a)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) / 2;
        const int c = flt_key_compare(t, list[mid].chr, list[mid].pos, a.chr, a.pos);
        if (c == 0) {  // the better score stays; on a tie the transcriptome alignment replaces what is there
            if (a.score < list[mid].score || (a.score == list[mid].score && a.is_transcriptome)) list[mid] = a;
            return n;
        }
        if (c < 0) lo = mid + 1;
        else hi = mid;
    }
    for (uint32_t k = n; k > lo; k--) list[k] = list[k - 1];
    list[lo] = a;
    return n + 1;
}

// ---- AlignmentFilter::Filter: classification of every (read 0 alignment, read 1 alignment) combination and the decision ------------
enum { FLT_NO_RC = 0, FLT_INTRAGENE = 1, FLT_INTRACHR = 2, FLT_INTERCHR = 3 };

struct FltPair {  // class AlignmentPair (AlignmentFilter.cpp:62-77): align1 = read 0's alignment, align2 = read 1's
    uint32_t a1, a2;  // indices into the two lists
    int32_t distance;
    uint32_t score;
#ifdef __CUDACC__
    __host__ __device__
#endif
    bool operator<(const FltPair &o) const { return score < o.score; }  // AlignmentFilter.cpp:95-97
};

// GTFGene::CheckBoundary with the default buffer of 1000 (SNAPLib/GTFReader.cpp:890-902); unsigned arithmetic as there
FLT_HD bool flt_check_boundary(const FltTables &t, int gene, int chr, uint32_t pos)
{
    if (t.g_chr[gene] != chr) return false;
    const uint32_t lo = t.g_start[gene] - 1000u + 1u;
    return pos >= (lo > 1u ? lo : 1u) && pos <= t.g_end[gene] + 1000u;
}

FLT_HD FltPair flt_make_pair(const FltAln &a1, const FltAln &a2, uint32_t i1, uint32_t i2)
{
    FltPair p;
    p.a1 = i1; p.a2 = i2; p.distance = 0;
    p.score = (uint32_t)(a1.score + a2.score);
    if (a1.direction && !a2.direction) p.distance = (int32_t)(a1.pos - a2.pos);
    else if (!a1.direction && a2.direction) p.distance = (int32_t)(a2.pos - a1.pos);
    return p;
}

// m1 = an alignment of read 0, m0 = an alignment of read 1 (the reference's names, AlignmentFilter.cpp:343-500)
FLT_HD int flt_classify(const FltTables &t, const FltAln &m0, const FltAln &m1)
{
    if ((m0.direction != 0) == (m1.direction != 0)) return FLT_NO_RC;
    if (m0.is_transcriptome && m1.is_transcriptome) {
        if (m0.chr != m1.chr) return FLT_INTERCHR;
        if (flt_check_boundary(t, m0.gene, m1.chr, m1.pos)) return FLT_INTRAGENE;
        if (flt_check_boundary(t, m1.gene, m0.chr, m0.pos)) return FLT_INTRAGENE;
        return FLT_INTRACHR;
    }
    if (m0.is_transcriptome) {
        if (m0.chr != m1.chr) return FLT_INTERCHR;
        return flt_check_boundary(t, m0.gene, m1.chr, m1.pos) ? FLT_INTRAGENE : FLT_INTRACHR;
    }
    if (m1.is_transcriptome) {
        if (m0.chr != m1.chr) return FLT_INTERCHR;
        return flt_check_boundary(t, m1.gene, m0.chr, m0.pos) ? FLT_INTRAGENE : FLT_INTRACHR;
    }
    return FLT_INTRAGENE;  // two genome alignments: "we can't be sure" (AlignmentFilter.cpp:462-464)
}

struct FltResult {  // the fields of PairedAlignmentResult the filter writes
    uint32_t location[2], tlocation[2];
    int32_t score[2], mapq[2];
    uint8_t status[2], direction[2], is_transcriptome[2];
    uint8_t aligned_as_pair;  // result->alignedAsPair as Filter leaves it: true only when a same-gene pair decided (AlignmentFilter.cpp:543-548)
    uint8_t pad;
};

// ProcessPairs (AlignmentFilter.cpp:1061-1180) once pairs[0] (and pairs[1]) are the elements std::sort leaves in front.
FLT_HD void flt_process_pairs(const FltTables &t, const FltAln *list0, const FltAln *list1, const FltPair *pairs, uint32_t n_pairs, uint32_t conf_diff,
                              uint32_t *genome_mapq, FltResult *r)
{
    const FltAln *al[2] = {&list0[pairs[0].a1], &list1[pairs[0].a2]};
    for (int e = 0; e < 2; e++) {
        if (al[e]->is_transcriptome) {
            r->tlocation[e] = al[e]->location;
            r->location[e] = t.piece_begin[al[e]->chr] + al[e]->pos - 1;  // Genome::getOffsetOfPiece(rname) + pos - 1
        } else {
            r->tlocation[e] = 0;
            r->location[e] = al[e]->location;
        }
        r->direction[e] = al[e]->direction;
        r->score[e] = al[e]->score;
        r->is_transcriptome[e] = al[e]->is_transcriptome;
    }
    if (!al[0]->is_transcriptome && !al[1]->is_transcriptome) *genome_mapq = (uint32_t)al[0]->mapq;
    bool unique = n_pairs == 1;
    if (!unique) unique = (uint32_t)(pairs[1].score - pairs[0].score) >= conf_diff;
    const uint32_t mq = *genome_mapq < 70u ? *genome_mapq : 70u;
    for (int e = 0; e < 2; e++) {
        r->mapq[e] = unique ? (int32_t)mq : 1;
        r->status[e] = unique ? 1 : 2;  // SingleHit : MultipleHits
    }
}

// CheckNoRC (AlignmentFilter.cpp:1039-1059)
FLT_HD void flt_check_no_rc(const FltAln *list0, const FltAln *list1, const FltPair *no_rc, uint32_t n, FltResult *r)
{
    for (uint32_t k = 0; k < n; k++) {
        if (list0[no_rc[k].a1].chr == list1[no_rc[k].a2].chr && no_rc[k].score < (uint32_t)(r->score[0] + r->score[1])) {
            r->status[0] = r->status[1] = 2;
            r->mapq[0] = r->mapq[1] = 1;
        }
    }
}

// FindPartialMatches (AlignmentFilter.cpp:957-1037) over the CharacterizeSeeds tuples of both reads (snapb200_characterize_batch
// layout: segment 2*read + direction, ascending (location, seed offset)).
FLT_HD void flt_partial_locations(const uint32_t *locs, const uint16_t *offs, uint64_t lo, uint64_t hi, bool rc, uint32_t read_len, uint32_t *out, uint32_t *n)
{
    uint64_t k = lo;
    while (k < hi) {
        uint64_t e = k;
        while (e + 1 < hi && locs[e + 1] == locs[k]) e++;
        out[(*n)++] = rc ? locs[k] + (read_len - offs[e]) : locs[k] + offs[k];  // smallest offset of the forward map, largest of the RC map
        k = e + 1;
    }
}

FLT_HD bool flt_partial_match(const FltTables &t, const uint32_t *l0, uint32_t n0, const uint32_t *l1, uint32_t n1, uint32_t max_spacing)
{
    for (uint32_t i = 0; i < n0; i++) {
        for (uint32_t j = 0; j < n1; j++) {
            const int p0 = flt_piece_at(t.piece_begin, (int)t.n_pieces, l0[i]), p1 = flt_piece_at(t.piece_begin, (int)t.n_pieces, l1[j]);
            if (p0 != p1 || p0 < 0) continue;  // before the first contig the reference dereferences a NULL piece; such locations never match here
            const int pos0 = (int)(l0[i] - t.piece_begin[p0] + 1), pos1 = (int)(l1[j] - t.piece_begin[p1] + 1);
            const uint32_t d = (uint32_t)(pos1 > pos0 ? pos1 - pos0 : pos0 - pos1);
            if (d < max_spacing) return true;
        }
    }
    return false;
}

// ---- std::sort as libstdc++ implements it (bits/stl_algo.h: introsort, threshold 16, median of three, final insertion sort) -----
// ProcessPairs sorts the pairs by score alone and takes the first two, so WHICH of several equal-scored pairs ends up in front is
// decided by this algorithm; the device version has to run the same one.  Mirrors GCC >= 4.9 (the toolchain the oracle is built with).
FLT_HD void flt_swap(FltPair &a, FltPair &b) { const FltPair t = a; a = b; b = t; }

FLT_HD void flt_adjust_heap(FltPair *first, long hole, long len, FltPair value)
{
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (first[child] < first[child - 1]) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    long parent = (hole - 1) / 2;  // __push_heap
    while (hole > top && first[parent] < value) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

FLT_HD void flt_heap_sort(FltPair *first, long len)  // __partial_sort(first, last, last): make_heap + sort_heap
{
    if (len >= 2) {
        for (long parent = (len - 2) / 2;; parent--) {
            flt_adjust_heap(first, parent, len, first[parent]);
            if (parent == 0) break;
        }
    }
    for (long last = len; last > 1;) {
        --last;
        const FltPair value = first[last];
        first[last] = first[0];
        flt_adjust_heap(first, 0, last, value);
    }
}

FLT_HD void flt_unguarded_linear_insert(FltPair *last)
{
    const FltPair val = *last;
    FltPair *next = last - 1;
    while (val < *next) { *last = *next; last = next; --next; }
    *last = val;
}

FLT_HD void flt_insertion_sort(FltPair *first, FltPair *last)
{
    if (first == last) return;
    for (FltPair *i = first + 1; i != last; ++i) {
        if (*i < *first) {
            const FltPair val = *i;
            for (FltPair *p = i; p != first; --p) *p = *(p - 1);
            *first = val;
        } else {
            flt_unguarded_linear_insert(i);
        }
    }
}

FLT_HD void flt_sort_pairs(FltPair *first, long n)
{
    if (n <= 0) return;
    // __introsort_loop with an explicit stack of (begin, end, depth) instead of the recursion on the right part
    long stack_b[64], stack_e[64];
    int stack_d[64], sp = 0;
    int depth = 0;
    for (long k = n; k > 1; k >>= 1) depth++;  // __lg(n)
    depth *= 2;
    long b = 0, e = n;
    for (;;) {
        while (e - b > 16) {
            if (depth == 0) { flt_heap_sort(first + b, e - b); break; }
            --depth;
            // __unguarded_partition_pivot
            const long mid = b + (e - b) / 2;
            FltPair &r = first[b], &x = first[b + 1], &y = first[mid], &z = first[e - 1];
            if (x < y) { if (y < z) flt_swap(r, y); else if (x < z) flt_swap(r, z); else flt_swap(r, x); }
            else if (x < z) flt_swap(r, x);
            else if (y < z) flt_swap(r, z);
            else flt_swap(r, y);
            long lo = b + 1, hi = e;
            for (;;) {
                while (first[lo] < first[b]) ++lo;
                --hi;
                while (first[b] < first[hi]) --hi;
                if (!(lo < hi)) break;
                flt_swap(first[lo], first[hi]);
                ++lo;
            }
            // recurse on [lo, e) first, as libstdc++ does, then continue with [b, lo): order does not change the result, the
            // two parts are disjoint -- so the right part goes on the stack and the loop goes on with the left one
            stack_b[sp] = lo; stack_e[sp] = e; stack_d[sp] = depth; sp++;
            e = lo;
        }
        if (sp == 0) break;
        sp--;
        b = stack_b[sp]; e = stack_e[sp]; depth = stack_d[sp];
    }
    // __final_insertion_sort
    if (n > 16) {
        flt_insertion_sort(first, first + 16);
        for (FltPair *i = first + 16; i != first + n; ++i) flt_unguarded_linear_insert(i);
    } else {
        flt_insertion_sort(first, first + n);
    }
}

// ---- AlignmentFilter::UnalignedRead (AlignmentFilter.cpp:742-933): the novel-splice search of a read without any alignment -------------
// Every distinct location of the read's seed maps becomes a partial alignment (the stretch of the read its seeds cover); every two of
// them that together cover the read and do not overlap are a candidate splice -- within reach of a gene that covers the first (dropped:
// the reference records nothing if there is any), on one chromosome, or on two.  The reference hands each to
// GTFReader::IntrachromosomalSplice / InterchromosomalSplice; here they are records the host appends to the same interval maps.
struct FltSeg { int32_t chr; uint32_t pos, pos_end, score; };

// GTFReader::IntervalGenes -> IntervalTree::findOverlapping (IntervalTree.h:192-211) finds a gene iff it overlaps [start, stop] (int
// comparisons) -- except that a tree of fewer than 64 intervals is a single UNSORTED node (:135-137, the root only sorts when it
// splits) whose scan is skipped altogether when stop < the start of the first interval in insertion order, i.e. of the first gene in
// gene-id order: gene_tree_min_stop is that start for n_genes < 64 and INT_MIN otherwise.
FLT_HD bool flt_gene_found(const FltTables &t, uint32_t g, int chr, uint32_t start, uint32_t stop)
{
    return t.g_chr[g] == chr && (int32_t)t.g_end[g] >= (int32_t)start && (int32_t)t.g_start[g] <= (int32_t)stop && (int32_t)stop >= t.gene_tree_min_stop;
}

enum { FLT_SPLICE_NONE = 0, FLT_SPLICE_GENE = 1, FLT_SPLICE_INTRACHR = 2, FLT_SPLICE_INTERCHR = 3 };

// the tests of the pair loop (:809-876) up to the gene query: NONE, INTERCHR, or "same chromosome" (returned as INTRACHR; the
// caller then asks flt_splice_in_gene)
FLT_HD int flt_splice_class(const FltSeg &a0, const FltSeg &a1, uint32_t read_len, uint32_t seed_len)
{
    if ((uint32_t)((int32_t)a0.score + (int32_t)a1.score) < read_len - seed_len) return FLT_SPLICE_NONE;  // int sum against unsigned, as there
    if (!(a0.pos > a1.pos_end) && !(a1.pos > a0.pos_end)) return FLT_SPLICE_NONE;
    return a0.chr != a1.chr ? FLT_SPLICE_INTERCHR : FLT_SPLICE_INTRACHR;
}

FLT_HD bool flt_splice_in_gene(const FltTables &t, const FltSeg &a0, const FltSeg &a1)
{
    for (uint32_t g = 0; g < t.n_genes; g++)
        if (flt_gene_found(t, g, a0.chr, a0.pos, a0.pos_end) && flt_check_boundary(t, (int)g, a1.chr, a1.pos)) return true;
    return false;
}

// The partial alignments in the reference's order (forward map ascending, then RC map ascending) from the CharacterizeSeeds tuples
// (snapb200_characterize_batch layout).  Returns their number, or -1 if there are more than cap.  A location before the first
// contig -- where the reference dereferences a NULL piece -- yields no segment.
FLT_HD int flt_unaligned_segments(const FltTables &t, const uint32_t *locs, const uint16_t *offs, uint64_t lo, uint64_t mid, uint64_t hi, uint32_t read_len,
                                  uint32_t seed_len, FltSeg *out, uint32_t cap)
{
    uint32_t n = 0;
    uint64_t k = lo;
    while (k < hi) {
        const bool rc = k >= mid;
        const uint64_t end = rc ? hi : mid;
        uint64_t e = k;
        while (e + 1 < end && locs[e + 1] == locs[k]) e++;
        const int p = flt_piece_at(t.piece_begin, (int)t.n_pieces, locs[k]);
        if (p >= 0) {
            if (n >= cap) return -1;
            const uint32_t length = (uint32_t)(offs[e] - offs[k]) + seed_len;  // (largest - smallest seed offset) + seedLen
            const int32_t pos0 = (int32_t)(locs[k] - t.piece_begin[p] + 1);
            const uint32_t start = rc ? (uint32_t)pos0 + read_len - ((uint32_t)offs[e] + seed_len) : (uint32_t)pos0 + offs[k];
            out[n].chr = p; out[n].pos = start; out[n].pos_end = start + length - 1; out[n].score = length;
            n++;
        }
        k = e + 1;
    }
    return (int)n;
}

struct FltSplice {  // one GTFReader::IntrachromosomalSplice (kind 2) / InterchromosomalSplice (kind 3) call
    uint32_t pair;      // pair index in the batch; the read id is read 0's or read 1's, whichever UnalignedRead was called for
    int32_t kind;
    int32_t chr[2];
    uint32_t pos[2], pos_end[2];
};

// The whole search for one read, serially (the specification the warp version of filter_warp.cuh is tested against).  out may be
// NULL (count only).  Returns the number of records, *kind = which call they are for.
FLT_HD uint64_t flt_unaligned_splices(const FltTables &t, const FltSeg *seg, uint32_t n, uint32_t read_len, uint32_t seed_len, uint32_t pair, int *kind,
                                      FltSplice *out)
{
    uint64_t c_intra = 0, c_inter = 0;
    bool gene = false;
    for (uint32_t i = 0; i < n && !gene; i++)
        for (uint32_t j = i + 1; j < n; j++) {
            const int c = flt_splice_class(seg[i], seg[j], read_len, seed_len);
            if (c == FLT_SPLICE_INTERCHR) c_inter++;
            else if (c == FLT_SPLICE_INTRACHR) { if (flt_splice_in_gene(t, seg[i], seg[j])) { gene = true; break; } c_intra++; }
        }
    *kind = gene ? FLT_SPLICE_NONE : c_intra ? FLT_SPLICE_INTRACHR : c_inter ? FLT_SPLICE_INTERCHR : FLT_SPLICE_NONE;
    if (*kind == FLT_SPLICE_NONE) return 0;
    const uint64_t total = *kind == FLT_SPLICE_INTRACHR ? c_intra : c_inter;
    if (out) {
        uint64_t w = 0;
        for (uint32_t i = 0; i < n; i++)
            for (uint32_t j = i + 1; j < n; j++) {
                int c = flt_splice_class(seg[i], seg[j], read_len, seed_len);
                if (c != *kind) continue;
                FltSplice &s = out[w++];
                s.pair = pair; s.kind = c;
                s.chr[0] = seg[i].chr; s.pos[0] = seg[i].pos; s.pos_end[0] = seg[i].pos_end;
                s.chr[1] = seg[j].chr; s.pos[1] = seg[j].pos; s.pos_end[1] = seg[j].pos_end;
            }
    }
    return total;
}

// ---- what the host still has to do per pair: the GTF statistics -------------------------------------------------------------------
// None of them feeds back into the pair's result, so the kernel reports WHAT to count and the host calls the reference's own public
// GTFReader methods in input order (AlignmentFilter.cpp:529-713): IncrementReadCount for a unique same-gene pair,
// IntrachromosomalPair / InterchromosomalPair for unique distant pairs, and UnalignedRead for a read without any alignment whose mate
// has some.
enum { FLT_EV_NONE = 0, FLT_EV_INCREMENT = 1, FLT_EV_INTRACHR = 2, FLT_EV_INTERCHR = 3 };
struct FltEvent {
    int32_t kind;
    int32_t unaligned;        // 0 none, 1: UnalignedRead(read 0), 2: UnalignedRead(read 1)
    int32_t transcript[2];    // of pairs[0].align1 / align2; -1 for a genome alignment (the reference passes "")
    int32_t chr[2];
    uint32_t pos_original[2], pos[2], pos_end[2];
};

// ---- the whole per-pair filter as one function over caller-provided scratch (no allocation, no recursion): what a kernel runs -------
struct FltParams { uint32_t max_dist, max_spacing, conf_diff; int32_t force_spacing; };

struct FltPairInput {
    uint32_t len[2];                       // data length of read 0 / read 1
    int32_t n_hits[2];                     // transcriptome multi-hits of each read (snapb200_single_multihit_batch)
    const uint32_t *hit_loc[2];
    const uint8_t *hit_rc[2];
    const int32_t *hit_score[2];
    uint32_t g_location[2];                // the genome pair (snapb200_paired_batch)
    int32_t g_score[2], g_mapq[2];
    uint8_t g_status[2], g_direction[2];
    const uint32_t *ch_loc[2];             // CharacterizeSeeds tuples of each read (snapb200_characterize_batch): locations, seed offsets,
    const uint16_t *ch_off[2];
    uint64_t ch_range[2][3];               // and the bounds of the forward and RC segment: [fwd begin, fwd end = rc begin, rc end)
};

struct FltScratch {
    FltAln *list[2]; uint32_t list_cap;    // room for list_cap + 1 entries each
    FltPair *pairs; uint32_t pair_cap;
    uint32_t *ploc[2]; uint32_t ploc_cap;
};

enum { FLT_OK = 0, FLT_SCRATCH_TOO_SMALL = 1 };

FLT_HD int flt_filter_pair(const FltTables &t, const FltParams &prm, const FltPairInput &in, const FltScratch &sc, FltResult *out, FltEvent *ev)
{
    uint32_t n[2] = {0, 0};
    FltAln a;
    for (int e = 0; e < 2; e++) {
        for (int k = 0; k < in.n_hits[e]; k++) {
            if (!flt_make_alignment(t, in.hit_loc[e][k], in.hit_rc[e][k] ? 1 : 0, in.hit_score[e][k], 0, true, in.len[e], prm.max_dist, &a)) continue;
            if (n[e] >= sc.list_cap) return FLT_SCRATCH_TOO_SMALL;
            n[e] = flt_insert(t, sc.list[e], n[e], a);
        }
    }
    for (int e = 0; e < 2; e++) {
        if (!flt_make_alignment(t, in.g_location[e], in.g_direction[e], in.g_score[e], in.g_mapq[e], false, in.len[e], prm.max_dist, &a)) continue;
        if (n[e] >= sc.list_cap) return FLT_SCRATCH_TOO_SMALL;
        n[e] = flt_insert(t, sc.list[e], n[e], a);
    }
    const FltAln *l0 = sc.list[0], *l1 = sc.list[1];
    FltResult r;
    for (int e = 0; e < 2; e++) {
        r.location[e] = in.g_location[e]; r.tlocation[e] = 0; r.score[e] = in.g_score[e]; r.mapq[e] = in.g_mapq[e];
        r.status[e] = in.g_status[e]; r.direction[e] = in.g_direction[e]; r.is_transcriptome[e] = 0;
    }
    r.aligned_as_pair = 0; r.pad = 0;
    ev->kind = FLT_EV_NONE;
    ev->unaligned = (n[0] == 0 && n[1] != 0) ? 1 : (n[1] == 0 && n[0] != 0) ? 2 : 0;
    for (int e = 0; e < 2; e++) { ev->transcript[e] = -1; ev->chr[e] = 0; ev->pos_original[e] = ev->pos[e] = ev->pos_end[e] = 0; }
    // which class decides: the first non-empty one in the order intragene, intrachromosomal, interchromosomal, no_rc
    uint32_t count[4] = {0, 0, 0, 0};
    for (uint32_t j = 0; j < n[1]; j++)
        for (uint32_t i = 0; i < n[0]; i++) count[flt_classify(t, l1[j], l0[i])]++;
    const int order[4] = {FLT_INTRAGENE, FLT_INTRACHR, FLT_INTERCHR, FLT_NO_RC};
    int chosen = -1;
    for (int k = 0; k < 4 && chosen < 0; k++) if (count[order[k]]) chosen = order[k];
    if (chosen < 0) {
        for (int e = 0; e < 2; e++) { r.location[e] = 0; r.tlocation[e] = 0; r.score[e] = 0; r.mapq[e] = 0; r.status[e] = 0; r.direction[e] = 0; r.is_transcriptome[e] = 0; }
    } else {
        if (count[chosen] > sc.pair_cap) return FLT_SCRATCH_TOO_SMALL;
        uint32_t np = 0;
        for (uint32_t j = 0; j < n[1]; j++)  // the reference's loop order: its outer map holds read 1's alignments
            for (uint32_t i = 0; i < n[0]; i++)
                if (flt_classify(t, l1[j], l0[i]) == chosen) sc.pairs[np++] = flt_make_pair(l0[i], l1[j], i, j);
        if (np > 1) flt_sort_pairs(sc.pairs, (long)np);
        uint32_t genome_mapq = 70;
        flt_process_pairs(t, l0, l1, sc.pairs, np, prm.conf_diff, &genome_mapq, &r);
        const FltPair p0 = sc.pairs[0];
        bool report = false;
        int kind = FLT_EV_NONE;
        if (chosen == FLT_INTRAGENE) {
            report = r.status[0] == 1;
            kind = FLT_EV_INCREMENT;
            r.aligned_as_pair = 1;
        } else {
            if (chosen != FLT_NO_RC && r.status[0] == 1 && count[FLT_NO_RC]) {  // CheckNoRC: any same-chromosome same-strand pair that scores better
                const uint32_t sum = (uint32_t)(r.score[0] + r.score[1]);
                bool hit = false;
                for (uint32_t j = 0; j < n[1] && !hit; j++)
                    for (uint32_t i = 0; i < n[0] && !hit; i++)
                        if (flt_classify(t, l1[j], l0[i]) == FLT_NO_RC && l0[i].chr == l1[j].chr && (uint32_t)(l0[i].score + l1[j].score) < sum) hit = true;
                if (hit) { r.status[0] = r.status[1] = 2; r.mapq[0] = r.mapq[1] = 1; }
            }
            const bool near = chosen == FLT_INTRACHR && (uint32_t)p0.distance <= prm.max_spacing;  // int against unsigned, as in the reference
            if (!near) {
                if (r.status[0] == 1) {  // FindPartialMatches
                    uint32_t c[2] = {0, 0};
                    for (int e = 0; e < 2; e++) {
                        if (in.ch_range[e][2] - in.ch_range[e][0] > sc.ploc_cap) return FLT_SCRATCH_TOO_SMALL;
                        flt_partial_locations(in.ch_loc[e], in.ch_off[e], in.ch_range[e][0], in.ch_range[e][1], false, in.len[e], sc.ploc[e], &c[e]);
                        flt_partial_locations(in.ch_loc[e], in.ch_off[e], in.ch_range[e][1], in.ch_range[e][2], true, in.len[e], sc.ploc[e], &c[e]);
                    }
                    if (flt_partial_match(t, sc.ploc[0], c[0], sc.ploc[1], c[1], prm.max_spacing)) { r.status[0] = r.status[1] = 2; r.mapq[0] = r.mapq[1] = 1; }
                }
                report = r.status[0] == 1;
                kind = chosen == FLT_INTRACHR ? FLT_EV_INTRACHR
                     : chosen == FLT_INTERCHR ? FLT_EV_INTERCHR
                     : (l0[p0.a1].chr == l1[p0.a2].chr ? FLT_EV_INTRACHR : FLT_EV_INTERCHR);
            }
        }
        if (report) {
            ev->kind = kind;
            const FltAln *al[2] = {&l0[p0.a1], &l1[p0.a2]};
            for (int e = 0; e < 2; e++) {
                ev->transcript[e] = al[e]->transcript; ev->chr[e] = al[e]->chr;
                ev->pos_original[e] = al[e]->pos_original; ev->pos[e] = al[e]->pos; ev->pos_end[e] = al[e]->pos_end;
            }
        }
    }
    // the run loop's epilogue (PairedAligner.cpp:646-663)
    if (prm.force_spacing && (r.status[0] == 1) != (r.status[1] == 1)) { r.status[0] = r.status[1] = 0; r.location[0] = r.location[1] = FLT_INVALID_LOC; }
    if (r.score[0] + r.score[1] >= 5) {
        if (r.mapq[0] < 50) r.mapq[0] /= 2;
        if (r.mapq[1] < 50) r.mapq[1] /= 2;
    }
    for (int e = 0; e < 2; e++) if (!r.is_transcriptome[e]) r.tlocation[e] = 0;
    *out = r;
    return FLT_OK;
}
