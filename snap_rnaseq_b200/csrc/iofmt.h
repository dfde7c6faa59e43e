// iofmt.h -- the per-record logic of the two I/O edges as plain functions:
//   * one FASTQ record: line splitting, starting-character checks, clipping   (FASTQReader::getNextRead,
//     SNAPLib/FASTQ.cpp:188-246, 253-297; Read::clip, SNAPLib/Read.h:357-404)
//   * one SAM line: every field of SAMFormat::writeRead except the CIGAR walk  (SNAPLib/SAM.cpp:803-1153,
//     SimpleReadWriter::writePair, SNAPLib/ReadWriter.cpp:132-217)
// The kernels of iokernels.cuh call these from device code; tests/hostsim compiles the same header with g++ to check
// the logic against the compiled reference on a box without a GPU (test infrastructure only -- the library has no
// host execution path for them).
#pragma once
#include <stdint.h>

#include "../../include/snapb200.h"
#include "filterfmt.h"  // FltTables: the exon / intron lists of the transcripts, for the CIGAR of a transcriptome alignment

#ifdef __CUDACC__
#define SNAP_HD __host__ __device__ __forceinline__
#else
#define SNAP_HD static inline
#endif

// ---- FASTQ ------------------------------------------------------------------------------------------------------
enum { FQ_OK = 0, FQ_BLANK_LINE = 1, FQ_BAD_START = 2, FQ_TOO_LONG = 3 };

struct FqRecord {
    uint32_t id_start, id_len;    // Read::getId / getIdLength (the '@' is not part of the id)
    uint32_t data_start, data_len;  // unclipped
    uint32_t qual_start;
    uint16_t front_clip, clipped_len;
    uint32_t end;                 // first byte after the record (FASTQ.cpp:236: one CR after the LF is skipped too)
    int error;
};

// FASTQReader::isValidStartingCharacterForNextLine[(i + 3) % 4][c] (FASTQ.cpp:253-297), i = line within the record
SNAP_HD bool fq_valid_start(int line, uint8_t c)
{
    switch (line) {
        case 0: return c == '@';
        case 1: return c == 'A' || c == 'C' || c == 'T' || c == 'G' || c == 'N' || c == 'a' || c == 'c' || c == 't' || c == 'g' || c == 'n';
        case 2: return c == '+';
        default: return c >= '!' && c <= '~';
    }
}

// Record r of a text whose '\n' positions are nl[] (ascending).  text[n_bytes] must be readable (any value but '\r'
// keeps the reference's behaviour at the end of its buffer, where a NUL follows).
SNAP_HD FqRecord fq_record(const uint8_t *text, uint64_t n_bytes, const uint32_t *nl, uint32_t r, int clipping)
{
    FqRecord rec;
    rec.error = FQ_OK;
    rec.id_start = rec.id_len = rec.data_start = rec.data_len = rec.qual_start = rec.end = 0;
    rec.front_clip = rec.clipped_len = 0;
    uint32_t start[4], len[4];
    for (int i = 0; i < 4; i++) {
        const uint32_t line = 4 * r + i;
        uint32_t s = 0;
        if (line > 0) {
            s = nl[line - 1] + 1;
            if (s < n_bytes && text[s] == '\r') s++;  // scan = newLine + (newLine[1] == '\r' ? 2 : 1)
        }
        const uint32_t e = nl[line];
        const uint32_t line_len = e - s;  // e >= s: a CR right after the previous LF cannot be an LF
        if (line_len == 0) { rec.error = FQ_BLANK_LINE; return rec; }
        if (!fq_valid_start(i, text[s])) { rec.error = FQ_BAD_START; return rec; }
        start[i] = s;
        len[i] = line_len - (text[e - 1] == '\r' ? 1 : 0);
    }
    const uint32_t e3 = nl[4 * r + 3];
    rec.end = e3 + 1;
    if (rec.end < n_bytes && text[rec.end] == '\r') rec.end++;
    rec.id_start = start[0] + 1;
    rec.id_len = len[0] - 1;
    rec.data_start = start[1];
    rec.data_len = len[1];
    rec.qual_start = start[3];  // the quality string is taken to be as long as the data (FASTQ.cpp:241)
    if (rec.data_len > 65535u) { rec.error = FQ_TOO_LONG; return rec; }
    // Read::clip (Read.h:357-404); clipping: 0 none, 1 front, 2 back, 3 both.  NoClipping is the state after init().
    uint32_t data_len = rec.data_len, front = 0;
    if (clipping != 0) {
        const uint8_t *q = text + rec.qual_start;
        if (clipping == 2 || clipping == 3) {
            while (data_len > 0 && q[data_len - 1] == '#') data_len--;
        }
        if (clipping == 1 || clipping == 3) {
            while (front < data_len && q[front] == '#') front++;
        }
        if (data_len - front < 50) {
            data_len = rec.data_len;
            front = 0;
        } else {
            data_len -= front;
        }
    }
    rec.front_clip = (uint16_t)front;
    rec.clipped_len = (uint16_t)data_len;
    return rec;
}

// FASTQReader::skipPartialRecord (FASTQ.cpp:113-184): where the first record begins in a buffer that may start in the middle
// of one -- the pattern {start of buffer | LF} '@' ... LF {ACTGNactg}* [CR] LF '+'.  Returns n_bytes when there is none.
// Bytes at and beyond n_bytes read as NUL (the reference's buffers end with one).
SNAP_HD uint64_t fq_record_start(const uint8_t *text, uint64_t n_bytes)
{
#define FQ_AT(i) ((i) < n_bytes ? text[(i)] : (uint8_t)0)
    uint64_t first = 0;
    if (FQ_AT(0) != '@') {
        while (first < n_bytes && text[first] != '\n' && text[first] != 0) first++;
        if (first >= n_bytes || text[first] != '\n') return n_bytes;
        first++;
    }
    for (;;) {
        if (first >= n_bytes) return n_bytes;
        uint64_t second = first;
        while (second < n_bytes && text[second] != '\n' && text[second] != 0) second++;
        if (second >= n_bytes || text[second] != '\n') return n_bytes;
        second++;
        if (text[first] != '@') { first = second; continue; }
        uint64_t third = second;
        for (;;) {
            const uint8_t c = FQ_AT(third);
            if (c == 'A' || c == 'C' || c == 'T' || c == 'G' || c == 'N' || c == 'a' || c == 'c' || c == 't' || c == 'g') third++;
            else break;
        }
        if (FQ_AT(third) == '\r') third++;
        if (FQ_AT(third) != '\n') { first = second; continue; }
        third++;
        if (FQ_AT(third) != '+') { first = second; continue; }
        return first;
    }
#undef FQ_AT
}

SNAP_HD uint8_t fq_upper(uint8_t c) { return (c >= 0x61 && c <= 0x7a) ? (uint8_t)(c - 0x20) : c; }  // TO_UPPER_CASE, Tables.cpp:74-81

// ---- SAM --------------------------------------------------------------------------------------------------------
#define SAMF_MULTI_SEGMENT 0x001
#define SAMF_ALL_ALIGNED 0x002
#define SAMF_UNMAPPED 0x004
#define SAMF_NEXT_UNMAPPED 0x008
#define SAMF_REVERSE_COMPLEMENT 0x010
#define SAMF_NEXT_REVERSED 0x020
#define SAMF_FIRST_SEGMENT 0x040
#define SAMF_LAST_SEGMENT 0x080

#define SAM_NAME_STAR (-1)   // "*"
#define SAM_NAME_EQUAL (-2)  // "="
#define SAM_INVALID_LOC 0xffffffffu
#define SAM_CIGAR_STRIDE 256  // >= 61 runs of at most 4 characters (k = MAX_K-1 = 30 edits) + NUL
#define SAM_SPLICED_CIGAR_STRIDE 512  // the same with soft clips and splice junctions inserted; a longer one is reported, not truncated

struct SamEnd {          // what getSAMData reads of one Read and its alignment
    uint32_t full_len;   // getUnclippedLength
    uint32_t front_clip; // getFrontClippedLength
    uint32_t clipped_len;  // getDataLength
    uint32_t location;   // already InvalidGenomeLocation when the status is NotFound (ReadWriter.cpp:103-105, 164-166)
    int direction;
    int mapq;
};

struct SamFields {
    int flags;
    int rname;  // piece index, SAM_NAME_STAR
    uint32_t pos;
    int mapq;
    int rnext;  // piece index, SAM_NAME_STAR, SAM_NAME_EQUAL
    uint32_t pnext;
    long long tlen;
    uint32_t clip_before, clip_after;
    int direction;  // FORWARD for an unmapped read (SAM.cpp:857-863)
    bool mapped;
    int ref_index, next_ref_index;  // pieceIndex / matePieceIndex as getSAMData leaves them (-1: none): BAM's refID / next_refID
    bool mate_mapped;
};

// Genome::getPieceAtLocation (Genome.cpp:357-374)
SNAP_HD int sam_piece_at(const uint32_t *piece_begin, int n_pieces, uint32_t location)
{
    int low = 0, high = n_pieces - 1;
    while (low <= high) {
        const int mid = (low + high) / 2;
        if (piece_begin[mid] <= location && (mid == n_pieces - 1 || piece_begin[mid + 1] > location)) return mid;
        if (piece_begin[mid] <= location) low = mid + 1;
        else high = mid - 1;
    }
    return 0;
}

// getSAMData (SAM.cpp:803-975) without the byte copies
SNAP_HD SamFields sam_fields(const uint32_t *piece_begin, int n_pieces, const SamEnd &me, bool has_mate, bool first_in_pair, const SamEnd &mate)
{
    SamFields f;
    f.flags = 0;
    f.rname = SAM_NAME_STAR;
    f.pos = 0;
    f.rnext = SAM_NAME_STAR;
    f.pnext = 0;
    f.tlen = 0;
    f.mapq = me.mapq;
    f.mapped = me.location != SAM_INVALID_LOC;
    f.ref_index = f.next_ref_index = -1;
    f.mate_mapped = false;
    f.direction = f.mapped ? me.direction : 0;
    if (f.direction == 1) {
        f.clip_before = me.full_len - me.clipped_len - me.front_clip;
        f.clip_after = me.front_clip;
    } else {
        f.clip_before = me.front_clip;
        f.clip_after = me.full_len - me.clipped_len - f.clip_before;
    }
    if (f.mapped) {
        if (f.direction == 1) f.flags |= SAMF_REVERSE_COMPLEMENT;
        f.rname = sam_piece_at(piece_begin, n_pieces, me.location);
        f.ref_index = f.rname;
        f.pos = me.location - piece_begin[f.rname] + 1;
        f.mapq = f.mapq < 0 ? 0 : (f.mapq > 70 ? 70 : f.mapq);
    } else {
        f.flags |= SAMF_UNMAPPED;
        f.mapq = 0;
    }
    if (has_mate) {
        f.flags |= SAMF_MULTI_SEGMENT;
        f.flags |= first_in_pair ? SAMF_FIRST_SEGMENT : SAMF_LAST_SEGMENT;
        if (mate.location != SAM_INVALID_LOC) {
            f.rnext = sam_piece_at(piece_begin, n_pieces, mate.location);
            f.next_ref_index = f.rnext;
            f.mate_mapped = true;
            f.pnext = mate.location - piece_begin[f.rnext] + 1;
            if (mate.direction == 1) f.flags |= SAMF_NEXT_REVERSED;
            if (!f.mapped) {
                f.ref_index = f.rnext;
                f.rname = f.rnext;
                f.rnext = SAM_NAME_EQUAL;
                f.pos = f.pnext;
            }
        } else {
            f.flags |= SAMF_NEXT_UNMAPPED;
            f.rnext = SAM_NAME_EQUAL;
            f.next_ref_index = f.ref_index;
            f.pnext = f.pos;
        }
        if (f.mapped && mate.location != SAM_INVALID_LOC) {
            f.flags |= SAMF_ALL_ALIGNED;
            // the reference mixes unsigned and _int64 here; the conversions are kept (SAM.cpp:949-962)
            const long long my_start = (long long)(uint32_t)(me.location - f.clip_before);
            const long long my_end = (long long)(uint32_t)(me.location + me.clipped_len + f.clip_after);
            const long long mate_before = (long long)mate.front_clip;
            const long long mate_after = (long long)(uint32_t)(mate.full_len - mate.clipped_len) - mate_before;
            const long long mate_start = (long long)mate.location - (mate.direction == 1 ? mate_after : mate_before);
            const long long mate_end = (long long)(uint32_t)(mate.location + mate.clipped_len) + (mate.direction == 0 ? mate_after : mate_before);
            if (f.rname == f.rnext) {
                if (my_start < mate_start) f.tlen = mate_end - my_start;
                else f.tlen = -(my_end - mate_start);
            }
        }
        if (f.rname >= 0 && f.rname == f.rnext) f.rnext = SAM_NAME_EQUAL;
    }
    return f;
}

// which end of pair p is written first, and the QNAME lengths (SimpleReadWriter::writePair, ReadWriter.cpp:147-168)
SNAP_HD int sam_pair_first(uint32_t loc0_masked, uint32_t loc1_masked) { return loc0_masked > loc1_masked ? 1 : 0; }

SNAP_HD bool sam_pair_trims_ids(const uint8_t *id0, uint32_t len0, const uint8_t *id1, uint32_t len1)
{
    if (len0 == len1 && len0 > 2 && id0[len0 - 2] == '/' && id1[len0 - 2] == '/') {
        const uint8_t c0 = id0[len0 - 1], c1 = id1[len1 - 1];
        if ((c0 == '1' || c0 == '2') && (c0 == '1' || c1 == '2') && c0 != c1) return true;  // the reference's condition, as written
    }
    return false;
}

SNAP_HD int sam_digits_u64(unsigned long long v)
{
    int n = 1;
    while (v > 0xffffffffull) { v /= 10; n++; }
    uint32_t w = (uint32_t)v;  // 64-bit division is a subroutine on the GPU; nearly every field fits 32 bits
    while (w >= 10) { w /= 10; n++; }
    return n;
}
SNAP_HD int sam_digits_i64(long long v)
{
    return v < 0 ? 1 + sam_digits_u64((unsigned long long)(-(v + 1)) + 1ull) : sam_digits_u64((unsigned long long)v);
}
SNAP_HD char *sam_put_u64(char *p, unsigned long long v)
{
    const int n = sam_digits_u64(v);
    int i = n - 1;
    for (; v > 0xffffffffull; i--) { p[i] = (char)('0' + (int)(v % 10)); v /= 10; }
    uint32_t w = (uint32_t)v;
    for (; i >= 0; i--) { p[i] = (char)('0' + (int)(w % 10)); w /= 10; }
    return p + n;
}
SNAP_HD char *sam_put_i64(char *p, long long v)
{
    if (v < 0) { *p++ = '-'; return sam_put_u64(p, (unsigned long long)(-(v + 1)) + 1ull); }
    return sam_put_u64(p, (unsigned long long)v);
}
SNAP_HD uint32_t sam_strlen(const char *s, uint32_t cap)
{
    uint32_t n = 0;
    while (n < cap && s[n]) n++;
    return n;
}

struct SamNames {  // piece names in HBM
    const char *blob;
    const uint32_t *off;  // [n_pieces + 1]
};
SNAP_HD uint32_t sam_name_len(const SamNames &nm, int id) { return id >= 0 ? nm.off[id + 1] - nm.off[id] : 1; }
SNAP_HD char *sam_put_name(char *p, const SamNames &nm, int id)
{
    if (id == SAM_NAME_STAR) { *p++ = '*'; return p; }
    if (id == SAM_NAME_EQUAL) { *p++ = '='; return p; }
    const uint32_t n = nm.off[id + 1] - nm.off[id];
    const char *s = nm.blob + nm.off[id];
    for (uint32_t i = 0; i < n; i++) p[i] = s[i];
    return p + n;
}

struct SamLine {  // what the measuring pass leaves for the writing pass
    uint32_t qname_len;  // after /1 /2 trimming and truncation at the first space (SAM.cpp:1075-1078)
    uint32_t seq_len;    // "%.*s" stops at a NUL: COMPLEMENT[] of a byte that is not ACGTNn is 0 (Tables.cpp:22-30)
    uint32_t qual_len;
    int32_t edit_distance;  // -1 unmapped / CIGAR "*"
    uint32_t cigar_len;     // of the LV string, without soft clips; 0 => "*"
    uint32_t spliced;       // a transcriptome alignment: the string is the whole field as insertSpliceJunctions leaves it (clips included;
                            // empty when the CIGAR could not be computed, SAM.cpp:1046-1061)
    uint32_t n_ops, ref_len;  // BAM records: the number of CIGAR operations and the reference bases they cover
};

// One "%d%c" of writeCigar (LandauVishkin.cpp:27-63): a count <= 0 writes nothing.  false: out of space.
SNAP_HD bool sam_cigar_put(char *out, uint32_t cap, uint32_t *n, int count, char op)
{
    if (count <= 0) return true;
    const int d = sam_digits_u64((unsigned long long)count);
    if (*n + (uint32_t)d + 1 > cap) return false;
    sam_put_u64(out + *n, (unsigned long long)count);
    out[*n + d] = op;
    *n += (uint32_t)d + 1;
    return true;
}

// LandauVishkinWithCigar::insertSpliceJunctions (SNAPLib/LandauVishkin.cpp:119-250) over GTFTranscript::Junctions
// (SNAPLib/GTFReader.cpp:1109-1139), for the tokens computeCigarString leaves (SAM.cpp:1206-1219: the soft clip before, the runs of the
// LV string `lv`, the soft clip after).  tr = the transcript of the transcriptome piece the alignment lies in, pos = its 1-based
// position in that transcript.  Kept as the reference does it, including: a junction is reported for an exon that ends exactly where
// the run ends (so a CIGAR can end in an N run); 'D' runs advance through the transcript, 'I' and 'S' do not; an intron of length
// <= 0 (abutting or overlapping exons) splits the run but prints nothing.  Returns the length written to out, -1 if cap is too small.
SNAP_HD int sam_splice_cigar(const FltTables &t, int tr, uint32_t pos, const char *lv, uint32_t lv_len, uint32_t clip_before, uint32_t clip_after,
                             char *out, uint32_t cap, uint32_t *n_calls = 0)
{
    uint32_t n = 0, prev = pos, current = pos, p = 0, calls = 0;
    int stage = 0;  // 0 the clip before, 1 the runs of lv, 2 the clip after
    for (;;) {
        uint32_t length;
        char op;
        if (stage == 0) {
            stage = 1;
            if (!clip_before) continue;
            length = clip_before; op = 'S';
        } else if (stage == 1) {
            if (p >= lv_len) { stage = 2; continue; }
            length = 0;
            while (p < lv_len && lv[p] >= '0' && lv[p] <= '9') length = length * 10 + (uint32_t)(lv[p++] - '0');
            op = p < lv_len ? lv[p++] : 'M';
        } else if (stage == 2) {
            stage = 3;
            if (!clip_after) continue;
            length = clip_after; op = 'S';
        } else {
            break;
        }
        if (op == 'I' || op == 'S') {
            calls++;
            if (!sam_cigar_put(out, cap, &n, (int)length, op)) return -1;
            continue;
        }
        current += length - 1;
        uint32_t remainder = length;
        if (tr >= 0) {
            const uint32_t query = prev, end_pos = prev + length;  // Junctions(prev, length): prev moves below, the query does not
            uint32_t cur_pos = 0;
            for (uint32_t k = t.t_feat_first[tr]; k < t.t_feat_first[tr + 1]; k++) {
                const uint32_t type = t.f_type[k], flen = t.f_end[k] - t.f_start[k] + 1;
                if (type == FLT_EXON) cur_pos += flen;
                if (query > cur_pos) continue;
                if (type == FLT_EXON) {
                    if (cur_pos >= end_pos) break;
                } else if (type == 2) {  // INTRON
                    const uint32_t first = cur_pos + 1;
                    if (first == pos) continue;
                    const int step = (int)(first - prev);
                    remainder -= (uint32_t)step;
                    if (step > 0) calls++;
                    if (step > 0 && !sam_cigar_put(out, cap, &n, step, op)) return -1;
                    calls++;
                    if (!sam_cigar_put(out, cap, &n, (int)flen, 'N')) return -1;
                    prev += (uint32_t)step;
                }
            }
        }
        if (remainder > 0) calls++;  // (the reference tests the unsigned remainder, then prints it as an int)
        if (!sam_cigar_put(out, cap, &n, (int)remainder, op)) return -1;  // the whole run when no junction was found
        current += 1;
        prev = current;
    }
    if (n_calls) *n_calls = calls;  // insertSpliceJunctions' return value: it also counts the runs that printed nothing
    return (int)n;
}

// The CIGAR field with soft clips (computeCigarString, SAM.cpp:1206-1222) -- length and bytes
SNAP_HD uint32_t sam_cigar_field_len(const SamFields &f, const SamLine &ln)
{
    if (f.mapped && ln.spliced) return ln.cigar_len;
    if (!f.mapped || ln.cigar_len == 0) return 1;
    uint32_t n = ln.cigar_len;
    if (f.clip_before > 0) n += sam_digits_u64(f.clip_before) + 1;
    if (f.clip_after > 0) n += sam_digits_u64(f.clip_after) + 1;
    return n;
}

SNAP_HD uint32_t sam_line_len(const SamFields &f, const SamLine &ln, const SamNames &nm, uint32_t rg_len)
{
    uint32_t n = ln.qname_len + 1;
    n += sam_digits_i64(f.flags) + 1;
    n += sam_name_len(nm, f.rname) + 1;
    n += sam_digits_u64(f.pos) + 1;
    n += sam_digits_i64(f.mapq) + 1;
    n += sam_cigar_field_len(f, ln) + 1;
    n += sam_name_len(nm, f.rnext) + 1;
    n += sam_digits_u64(f.pnext) + 1;
    n += sam_digits_i64(f.tlen) + 1;
    n += ln.seq_len + 1;
    n += ln.qual_len;
    if (rg_len) n += 6 + rg_len;                         // "\tRG:Z:" + group
    n += 10;                                             // "\tPG:Z:SNAP"
    n += 6 + sam_digits_i64(ln.edit_distance) + 1;       // "\tNM:i:%d" + "\n"
    return n;
}

// Everything before SEQ; returns the position where SEQ starts.
SNAP_HD char *sam_put_prefix(char *p, const uint8_t *id, const SamFields &f, const SamLine &ln, const SamNames &nm, const char *cigar)
{
    for (uint32_t i = 0; i < ln.qname_len; i++) p[i] = (char)id[i];
    p += ln.qname_len;
    *p++ = '\t';
    p = sam_put_i64(p, f.flags); *p++ = '\t';
    p = sam_put_name(p, nm, f.rname); *p++ = '\t';
    p = sam_put_u64(p, f.pos); *p++ = '\t';
    p = sam_put_i64(p, f.mapq); *p++ = '\t';
    if (f.mapped && ln.spliced) {
        for (uint32_t i = 0; i < ln.cigar_len; i++) p[i] = cigar[i];
        p += ln.cigar_len;
    } else if (!f.mapped || ln.cigar_len == 0) {
        *p++ = '*';
    } else {
        if (f.clip_before > 0) { p = sam_put_u64(p, f.clip_before); *p++ = 'S'; }
        for (uint32_t i = 0; i < ln.cigar_len; i++) p[i] = cigar[i];
        p += ln.cigar_len;
        if (f.clip_after > 0) { p = sam_put_u64(p, f.clip_after); *p++ = 'S'; }
    }
    *p++ = '\t';
    p = sam_put_name(p, nm, f.rnext); *p++ = '\t';
    p = sam_put_u64(p, f.pnext); *p++ = '\t';
    p = sam_put_i64(p, f.tlen); *p++ = '\t';
    return p;
}

// SEQ, a tab and QUAL, strided over `nlanes` cooperating callers (SAM.cpp:866-885); p = start of SEQ
SNAP_HD void sam_put_seq_qual(char *p, const uint8_t *bases, const uint8_t *quals, uint32_t full_len, int direction, const SamLine &ln, uint32_t lane,
                              uint32_t nlanes)
{
    for (uint32_t i = lane; i < ln.seq_len; i += nlanes) {
        uint8_t c;
        if (direction == 1) {
            const uint8_t b = bases[full_len - 1 - i];
            c = b == 'A' ? 'T' : b == 'C' ? 'G' : b == 'G' ? 'C' : b == 'T' ? 'A' : b == 'N' ? 'N' : b == 'n' ? 'n' : 0;
        } else {
            c = bases[i];
        }
        p[i] = (char)c;
    }
    if (lane == 0) p[ln.seq_len] = '\t';
    char *q = p + ln.seq_len + 1;
    for (uint32_t i = lane; i < ln.qual_len; i += nlanes) q[i] = (char)(direction == 1 ? quals[full_len - 1 - i] : quals[i]);
}

// Everything after QUAL; p = first byte after QUAL; returns the end of the line
SNAP_HD char *sam_put_suffix(char *p, const SamLine &ln, const char *rg, uint32_t rg_len)
{
    if (rg_len) {
        const char t[6] = {'\t', 'R', 'G', ':', 'Z', ':'};
        for (int i = 0; i < 6; i++) *p++ = t[i];
        for (uint32_t i = 0; i < rg_len; i++) *p++ = rg[i];
    }
    const char t2[16] = {'\t', 'P', 'G', ':', 'Z', ':', 'S', 'N', 'A', 'P', '\t', 'N', 'M', ':', 'i', ':'};
    for (int i = 0; i < 16; i++) *p++ = t2[i];
    p = sam_put_i64(p, ln.edit_distance);
    *p++ = '\n';
    return p;
}

// ---- which read a SAM line is ------------------------------------------------------------------------------------------
struct SamReadsDev {  // a snapb200_sam_reads (arrays resident where the caller runs)
    const uint32_t *offsets;
    const uint8_t *bases, *quals;
    const uint16_t *front_clip, *clipped_len;
    const uint32_t *id_offsets;
    const uint8_t *ids;
};

struct SamInputs {
    SamReadsDev rd[2];
    const snapb200_sam_alignment *aln[2];
    int paired;
};

struct SamWho {  // which read a line is, and its mate
    int e;  // 0/1: which batch
    uint32_t i;
    bool has_mate, first_in_pair;
    SamEnd me, mate;
    bool skip;
};

SNAP_HD SamEnd sam_end_of(const SamInputs &a, int e, uint32_t i)
{
    const snapb200_sam_alignment al = a.aln[e][i];
    SamEnd s;
    s.full_len = a.rd[e].offsets[i + 1] - a.rd[e].offsets[i];
    s.front_clip = a.rd[e].front_clip[i];
    s.clipped_len = a.rd[e].clipped_len[i];
    s.location = al.status == SNAPB200_NOT_FOUND ? SAM_INVALID_LOC : al.location;  // ReadWriter.cpp:103-105, 164-166
    s.direction = al.direction;
    s.mapq = al.mapq;
    return s;
}

// single-end: line i is read i.  Pairs: lines 2p, 2p+1 are the ends of pair p in the order writePair writes them; the one
// written first is flagged SAM_FIRST_SEGMENT (ReadWriter.cpp:167-189).
SNAP_HD SamWho sam_who(const SamInputs &a, uint32_t line)
{
    SamWho w;
    if (!a.paired) {
        w.e = 0; w.i = line; w.has_mate = false; w.first_in_pair = false;
        w.me = sam_end_of(a, 0, line);
        w.mate = w.me;
        w.skip = a.aln[0][line].skip != 0;
        return w;
    }
    const uint32_t p = line >> 1;
    const SamEnd e0 = sam_end_of(a, 0, p), e1 = sam_end_of(a, 1, p);
    const int first = sam_pair_first(e0.location, e1.location);
    w.first_in_pair = (line & 1) == 0;
    w.e = w.first_in_pair ? first : 1 - first;
    w.i = p;
    w.has_mate = true;
    w.me = w.e ? e1 : e0;
    w.mate = w.e ? e0 : e1;
    w.skip = a.aln[w.e][p].skip != 0;
    return w;
}

// ---- BAM records (BAMFormat::writeRead, SNAPLib/Bam.cpp:596-790) ----------------------------------------------------------------
// The same fields in binary: the 36-byte BAMAlignment head, NUL-terminated name, CIGAR operations (count << 4 | code of "MIDNSHP=X"),
// bases as nibbles of "=ACMGRSVTWYHKDBN", qualities minus '!', then the optional fields RG:Z (if there is a read group), PG:Z:SNAP and
// NM:i (int32).  Unlike the SAM writer this one does not cut the name at a space, and a CIGAR that could not be computed is zero
// operations.  NM of an unmapped read is an uninitialised variable in the reference (Bam.cpp:644); -1 is written here.
SNAP_HD int bam_reg2bin(int beg, int end)
{  // BAMAlignment::reg2bin, Bam.cpp:278-291
    --end;
    if (beg >> 14 == end >> 14) return ((1 << 15) - 1) / 7 + (beg >> 14);
    if (beg >> 17 == end >> 17) return ((1 << 12) - 1) / 7 + (beg >> 17);
    if (beg >> 20 == end >> 20) return ((1 << 9) - 1) / 7 + (beg >> 20);
    if (beg >> 23 == end >> 23) return ((1 << 6) - 1) / 7 + (beg >> 23);
    if (beg >> 26 == end >> 26) return ((1 << 3) - 1) / 7 + (beg >> 26);
    return 0;
}
SNAP_HD uint32_t bam_seq_code(uint8_t c)
{  // BAMAlignment::SeqToCode (Bam.cpp:266-271): position in "=ACMGRSVTWYHKDBN", 0 for everything else
    switch (c) {
        case 'A': return 1; case 'C': return 2; case 'M': return 3; case 'G': return 4; case 'R': return 5; case 'S': return 6; case 'V': return 7;
        case 'T': return 8; case 'W': return 9; case 'Y': return 10; case 'H': return 11; case 'K': return 12; case 'D': return 13; case 'B': return 14;
        case 'N': return 15; default: return 0;
    }
}
SNAP_HD uint32_t bam_cigar_code(char op)
{  // BAMAlignment::CigarToCode: position in "MIDNSHP=X"
    switch (op) { case 'I': return 1; case 'D': return 2; case 'N': return 3; case 'S': return 4; case 'H': return 5; case 'P': return 6; case '=': return 7; case 'X': return 8; default: return 0; }
}
// The runs of a CIGAR string as BAM operations: counts them and the reference bases they cover (CigarCodeToRefBase: M D N P = X),
// and writes them to out (4 bytes each, any alignment) unless out is NULL.
SNAP_HD void bam_cigar_ops(const char *s, uint32_t len, uint32_t *n_ops, uint32_t *ref_len, uint8_t *out)
{
    uint32_t p = 0, n = 0, ref = 0;
    while (p < len) {
        uint32_t count = 0;
        while (p < len && s[p] >= '0' && s[p] <= '9') count = count * 10 + (uint32_t)(s[p++] - '0');
        const char op = p < len ? s[p++] : 'M';
        const uint32_t code = bam_cigar_code(op);
        if (code != 1 && code != 4 && code != 5) ref += count;
        if (out) {
            const uint32_t v = (count << 4) | code;
            out[4 * n] = (uint8_t)v; out[4 * n + 1] = (uint8_t)(v >> 8); out[4 * n + 2] = (uint8_t)(v >> 16); out[4 * n + 3] = (uint8_t)(v >> 24);
        }
        n++;
    }
    *n_ops = n;
    *ref_len = ref;
}
// what computeCigarOps / insertSpliceJunctions return: the operations of the string plus one for each soft clip of a genome alignment
SNAP_HD void bam_count_ops(const SamFields &f, SamLine *ln, const char *cigar)
{
    ln->n_ops = ln->ref_len = 0;
    if (!f.mapped || ln->cigar_len == 0) return;
    bam_cigar_ops(cigar, ln->cigar_len, &ln->n_ops, &ln->ref_len, (uint8_t *)0);
    if (!ln->spliced) ln->n_ops += (f.clip_before > 0) + (f.clip_after > 0);
}
SNAP_HD uint32_t bam_record_len(const SamLine &ln, uint32_t full_len, uint32_t rg_len)
{
    return 36 + ln.qname_len + 1 + 4 * ln.n_ops + (full_len + 1) / 2 + full_len + (rg_len ? 4 + rg_len : 0) + 8 + 7;  // Bam.cpp:709-714
}
SNAP_HD void bam_put_u32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
SNAP_HD void bam_put_u16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
// Everything before the bases: head, name, CIGAR operations.  Returns where the bases start.
SNAP_HD uint8_t *bam_put_head(uint8_t *p, uint32_t record_len, const uint8_t *id, const SamFields &f, const SamLine &ln, uint32_t full_len, const char *cigar)
{
    bam_put_u32(p, record_len - 4);
    bam_put_u32(p + 4, (uint32_t)f.ref_index);
    bam_put_u32(p + 8, f.pos - 1);
    p[12] = (uint8_t)(ln.qname_len + 1);
    p[13] = (uint8_t)f.mapq;
    const int ref_len = ln.n_ops > 0 ? (int)ln.ref_len : (int)full_len;
    const int bin = f.mapped ? bam_reg2bin((int)f.pos - 1, (int)f.pos - 1 + ref_len)
                             : (f.mate_mapped ? bam_reg2bin((int)f.pnext - 1, (int)f.pnext) : bam_reg2bin(-1, 0));
    bam_put_u16(p + 14, (uint32_t)bin);
    bam_put_u16(p + 16, ln.n_ops);
    bam_put_u16(p + 18, (uint32_t)f.flags);
    bam_put_u32(p + 20, full_len);
    bam_put_u32(p + 24, (uint32_t)f.next_ref_index);
    bam_put_u32(p + 28, f.pnext - 1);
    bam_put_u32(p + 32, (uint32_t)(int)f.tlen);
    p += 36;
    for (uint32_t i = 0; i < ln.qname_len; i++) p[i] = id[i];
    p[ln.qname_len] = 0;
    p += ln.qname_len + 1;
    if (ln.n_ops) {
        uint32_t k = 0, n, r;
        if (!ln.spliced && f.clip_before > 0) { bam_put_u32(p, (f.clip_before << 4) | 4u); k = 1; }
        bam_cigar_ops(cigar, ln.cigar_len, &n, &r, p + 4 * k);
        k += n;
        if (!ln.spliced && f.clip_after > 0) { bam_put_u32(p + 4 * k, (f.clip_after << 4) | 4u); k++; }
        p += 4 * k;
    }
    return p;
}
// bases and qualities, strided over nlanes cooperating callers; p = where the bases start
SNAP_HD void bam_put_seq_qual(uint8_t *p, const uint8_t *bases, const uint8_t *quals, uint32_t full_len, int direction, uint32_t lane, uint32_t nlanes)
{
    const uint32_t nb = (full_len + 1) / 2;
    for (uint32_t j = lane; j < nb; j += nlanes) {
        uint32_t v = 0;
        for (uint32_t h = 0; h < 2; h++) {
            const uint32_t i = 2 * j + h;
            uint32_t code = 0;
            if (i < full_len) {
                uint8_t c;
                if (direction == 1) {
                    const uint8_t b = bases[full_len - 1 - i];
                    c = b == 'A' ? 'T' : b == 'C' ? 'G' : b == 'G' ? 'C' : b == 'T' ? 'A' : b == 'N' ? 'N' : b == 'n' ? 'n' : 0;
                } else {
                    c = bases[i];
                }
                code = bam_seq_code(c);
            }
            v = (v << 4) | code;
        }
        p[j] = (uint8_t)v;
    }
    uint8_t *q = p + nb;
    for (uint32_t i = lane; i < full_len; i += nlanes) q[i] = (uint8_t)((direction == 1 ? quals[full_len - 1 - i] : quals[i]) - '!');
}
// the optional fields; p = first byte after the qualities
SNAP_HD uint8_t *bam_put_aux(uint8_t *p, const SamLine &ln, const char *rg, uint32_t rg_len)
{
    if (rg_len) {
        p[0] = 'R'; p[1] = 'G'; p[2] = 'Z';
        for (uint32_t i = 0; i < rg_len; i++) p[3 + i] = (uint8_t)rg[i];
        p[3 + rg_len] = 0;
        p += 4 + rg_len;
    }
    const uint8_t pg[8] = {'P', 'G', 'Z', 'S', 'N', 'A', 'P', 0};
    for (int i = 0; i < 8; i++) p[i] = pg[i];
    p += 8;
    p[0] = 'N'; p[1] = 'M'; p[2] = 'i';
    bam_put_u32(p + 3, (uint32_t)ln.edit_distance);
    return p + 7;
}
