// filterfmt.h -- first pieces of the device AlignmentFilter (NOT a product path yet: nothing in the
// library includes this file).  Plain functions over flat tables, written the way iofmt.h is: the kernels of the next round will
// call them, and tests/hostsim runs them on the host against the compiled reference today.
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
