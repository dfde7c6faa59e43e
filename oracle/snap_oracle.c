/*
 * snap_oracle.c -- CPU restatement of the SNAP-RNA read-alignment hot path.
 *
 * TEST INFRASTRUCTURE ONLY ("port" oracle).  Plain sequential C, one function per reference function,
 * each citing the reference file:line it restates (paths relative to andrewmagis/snap-rnaseq).  It is
 * pinned against (a) the reference's own known-answer tests (tests/LandauVishkinTest.cpp) and (b) the
 * reference itself compiled into oracle/_ref/libsnapref.so -- see tests/test_oracle_*.py.  Nothing in the
 * product (snap_rnaseq_b200/) includes, links or calls this file.
 *
 * Results use the structs of include/snapb200.h so oracle, reference and CUDA outputs compare field by
 * field.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/snapb200.h"

#define MAXK SNAPB200_MAX_K /* 31 */
#define INVALID_LOC 0xffffffffu
#define UNUSED_SCORE 0xffffu
#define BUCKET 48u /* BaseAligner::maxMergeDist == hashTableElementSize, BaseAligner.h:163,196 */

/* ------------------------------------------------------------------------------------------------ */
/* probability tables: LandauVishkin.cpp:601-653                                                     */
/* ------------------------------------------------------------------------------------------------ */
#define MAX_INDELS 10000
static double g_phred[256];
static double g_indel[MAX_INDELS + 1];
static double g_perfect[SNAPB200_MAX_READ_LENGTH + 1];
static int g_tables_ready = 0;

static void init_tables(void)
{
    if (g_tables_ready) return;
    const double snp = 0.001, gap_open = 0.001, gap_extend = 0.5; /* BaseAligner.h:264-266 */
    g_indel[0] = 1.0;
    g_indel[1] = gap_open;
    for (int i = 2; i <= MAX_INDELS; i++) g_indel[i] = g_indel[i - 1] * gap_extend;
    for (int i = 0; i < 256; i++) g_phred[i] = snp;
    for (int i = 33; i <= 93 + 33; i++) g_phred[i] = 1.0 - (1.0 - pow(10.0, -1.0 * (i - 33.0) / 10.0)) * (1.0 - snp);
    g_perfect[0] = 1.0;
    for (int i = 1; i <= SNAPB200_MAX_READ_LENGTH; i++) g_perfect[i] = g_perfect[i - 1] * (1 - snp);
    g_tables_ready = 1;
}

/* pow(double,int) as the reference's C++98 build evaluates it: libstdc++'s std::pow(double,int) overload is
 * __builtin_powi, i.e. libgcc's __powidf2 square-and-multiply, NOT libm's pow (BaseAligner.cpp:1227,
 * IntersectingPairedEndAligner.cpp:834 pass an int seedLen).  Restated here so the result does not depend
 * on how this file is compiled. */
static double powi_ref(double x, int m)
{
    unsigned n = m < 0 ? -(unsigned)m : (unsigned)m;
    double y = (n % 2) ? x : 1;
    while (n >>= 1) {
        x = x * x;
        if (n % 2) y *= x;
    }
    return m < 0 ? 1 / y : y;
}

/* the three tables, for upload by tests that want to compare them with the product's host-built copies */
int oracle_prob_tables(double *phred256, double *indel64, double *perfect501)
{
    init_tables();
    memcpy(phred256, g_phred, sizeof(g_phred));
    memcpy(indel64, g_indel, 64 * sizeof(double));
    memcpy(perfect501, g_perfect, sizeof(g_perfect));
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* computeMAPQ: mapq.h:32-65                                                                         */
/* ------------------------------------------------------------------------------------------------ */
static int compute_mapq(double p_all, double p_best, int score, int popular_skipped)
{
    if (!(p_all > p_best)) p_all = p_best; /* __max(all, best) */
    if (p_all == p_best && popular_skipped == 0 && score < 5) return 70;
    double correct = p_best / p_all;
    int base;
    if (correct >= 1) {
        base = 69;
    } else {
        int v = (int)(-10 * log10(1 - correct));
        base = v < 69 ? v : 69;
    }
    int pen = popular_skipped - 10;
    if (pen < 0) pen = 0;
    base -= pen / 2;
    return base > 0 ? base : 0;
}

int oracle_mapq_batch(uint32_t n, const double *p_all, const double *p_best, const int32_t *score,
                      const int32_t *popular, int32_t *mapq)
{
    for (uint32_t i = 0; i < n; i++) mapq[i] = compute_mapq(p_all[i], p_best[i], score[i], popular[i]);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* Landau-Vishkin.  Strings are accessed through bounds-checked getters: a byte outside the pattern    */
/* is 0x00 and a byte outside the text is 0x01, so out-of-range peeks never match (the reference peeks */
/* with 8-byte loads and clamps afterwards, LandauVishkin.h:325-354).                                 */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    const uint8_t *p; /* pattern[0..plen) */
    int plen;
    const uint8_t *t; /* text: forward text[i] = t[i]; backward text[i] = t[-1-i] (t points one past) */
    int tlen;         /* bytes the caller declared                                                   */
    int dir;          /* +1 / -1                                                                     */
    int t_lo, t_hi;   /* readable window of t in "text index" space: indices in [t_lo, t_hi) are real */
} lv_strings;

static inline int pat_at(const lv_strings *s, int i) { return (i >= 0 && i < s->plen) ? s->p[i] : 0x00; }
static inline int txt_at(const lv_strings *s, int i)
{
    if (i < s->t_lo || i >= s->t_hi) return 0x01;
    return s->dir > 0 ? s->t[i] : s->t[-1 - i];
}

/* length of the common run pattern[pi..] vs text[ti..], clamped so that pi+run <= end */
static int match_run(const lv_strings *s, int pi, int ti, int end)
{
    int n = 0;
    while (pi + n < end && pat_at(s, pi + n) == txt_at(s, ti + n)) n++;
    return n;
}

/* the 0,+1,-1,+2,-2,... diagonal order of LandauVishkin.h:180-182,311 */
static inline int next_d_score(int d) { return d > 0 ? -d : -d + 1; }
/* the 0,-1,+1,-2,+2,... order of LandauVishkin.cpp:313 */
static inline int next_d_cigar(int d) { return d >= 0 ? -(d + 1) : -d; }

#define LROW (2 * MAXK + 1)
#define LAT(L, e, d) (L)[(e) * LROW + MAXK + (d)]

/* LandauVishkin<TEXT_DIRECTION>::computeEditDistance, LandauVishkin.h:211-455 */
static int lv_score(const lv_strings *s, const uint8_t *qual, int k, double *match_prob, int *net_indel)
{
    int L[(MAXK + 1) * LROW];
    char A[(MAXK + 1) * LROW];
    for (int i = 0; i < (MAXK + 1) * LROW; i++) L[i] = -2; /* constructor, LandauVishkin.h:168 */
    int local_indel;
    if (!net_indel) net_indel = &local_indel;
    *net_indel = 0;
    if (k > MAXK - 1) k = MAXK - 1; /* :234 */
    if (match_prob) *match_prob = 1.0;
    const int plen = s->plen, tlen = s->tlen;
    int end = plen < tlen ? plen : tlen;
    LAT(L, 0, 0) = match_run(s, 0, 0, end); /* :264-288 */
    if (LAT(L, 0, 0) == end) {              /* :290-305 */
        int result = plen > end ? plen - end : 0;
        if (match_prob) *match_prob = g_perfect[plen];
        return result > k ? -1 : result;
    }
    for (int e = 1; e <= k; e++) {
        for (int d = 0; d != e + 1; d = next_d_score(d)) {
            int best = LAT(L, e - 1, d) + 1; /* substitution */
            char act = 'X';
            int left = LAT(L, e - 1, d - 1);
            if (left > best) { best = left; act = 'D'; }
            int right = LAT(L, e - 1, d + 1) + 1;
            if (right > best) { best = right; act = 'I'; }
            LAT(A, e, d) = act;
            /* :325-354: extend only if the byte AT best matches (peek, possibly out of range), then clamp */
            if (pat_at(s, best) == txt_at(s, d + best)) {
                int dend = plen < tlen - d ? plen : tlen - d;
                if (best < dend) {
                    best += match_run(s, best, d + best, dend);
                } else {
                    best = dend; /* the 8-byte loop clamps to end even when it started beyond it */
                }
            }
            if (best == plen) { /* :356 */
                if (match_prob) {
                    /* backtrace, :379-431 */
                    char bt_act[MAXK + 1];
                    int bt_matched[MAXK + 1];
                    double prob = 1.0;
                    int cur_d = d;
                    /* NB the reference reads L[e][d] here BEFORE storing it (:385-391 vs :447), i.e. a stale cell;
                     * it only reaches backtraceMatched[e], which no output depends on.  `best` is used instead. */
                    for (int ce = e; ce >= 1; ce--) {
                        char a = LAT(A, ce, cur_d);
                        int here = (ce == e) ? best : LAT(L, ce, cur_d);
                        bt_act[ce] = a;
                        if (a == 'I') {
                            bt_matched[ce] = here - LAT(L, ce - 1, cur_d + 1) - 1;
                            cur_d = cur_d + 1;
                        } else if (a == 'D') {
                            bt_matched[ce] = here - LAT(L, ce - 1, cur_d - 1);
                            cur_d = cur_d - 1;
                        } else {
                            bt_matched[ce] = here - LAT(L, ce - 1, cur_d) - 1;
                        }
                    }
                    int ce = 1;
                    int offset = LAT(L, 0, 0);
                    while (ce <= e) {
                        char a = bt_act[ce];
                        int count = 1;
                        while (ce + 1 <= e && bt_matched[ce] == 0 && bt_act[ce + 1] == a) { count++; ce++; }
                        if (a == 'I') {
                            prob *= g_indel[count];
                            offset += count;
                            *net_indel += count;
                        } else if (a == 'D') {
                            prob *= g_indel[count];
                            offset -= count;
                            *net_indel -= count;
                        } else {
                            for (int i = 0; i < count; i++) {
                                int qi = offset < 0 ? 0 : offset;
                                if (qi > plen - 1) qi = plen - 1;
                                prob *= g_phred[qual[qi]];
                                offset++;
                            }
                        }
                        offset += bt_matched[ce];
                        ce++;
                    }
                    prob *= g_perfect[plen - e];
                    *match_prob = prob;
                }
                return e;
            }
            LAT(L, e, d) = best;
        }
    }
    return -1;
}

/* COMPACT_CIGAR_STRING writer: writeCigar, LandauVishkin.cpp:27-64.  Returns 0 if it did not fit. */
typedef struct { char *buf; int left; } cigar_out;
static int cigar_put(cigar_out *o, int count, char code)
{
    if (count <= 0) return 1;
    if (o->left == 0) { o->buf[-1] = 0; return 0; }
    int w = snprintf(o->buf, o->left, "%d%c", count, code);
    if (w > o->left - 1) return 0;
    o->buf += w;
    o->left -= w;
    return 1;
}

/* LandauVishkinWithCigar::computeEditDistance, LandauVishkin.cpp:252-535 (COMPACT_CIGAR_STRING) */
static int lv_cigar(const lv_strings *s, int k, char *cigar, int cigar_len, int use_m)
{
    int L[(MAXK + 1) * LROW];
    char A[(MAXK + 1) * LROW];
    for (int i = 0; i < (MAXK + 1) * LROW; i++) L[i] = -2;
    cigar_out o = {cigar, cigar_len};
    const int plen = s->plen, tlen = s->tlen;
    int end = plen < tlen ? plen : tlen;
    LAT(L, 0, 0) = match_run(s, 0, 0, end);
    if (LAT(L, 0, 0) == end) { /* :284-306 */
        if (use_m) {
            if (!cigar_put(&o, plen, 'M')) return -2;
        } else {
            if (!cigar_put(&o, end, '=')) return -2;
            if (plen > end && !cigar_put(&o, plen - end, 'X')) return -2;
        }
        return 0;
    }
    for (int e = 1; e <= k; e++) {
        for (int d = 0; d != -(e + 1); d = next_d_cigar(d)) {
            int best = LAT(L, e - 1, d) + 1;
            char act = 'X';
            int left = LAT(L, e - 1, d - 1);
            if (left > best) { best = left; act = 'D'; }
            int right = LAT(L, e - 1, d + 1) + 1;
            if (right > best) { best = right; act = 'I'; }
            LAT(A, e, d) = act;
            if (pat_at(s, best) == txt_at(s, d + best)) {
                int dend = plen < tlen - d ? plen : tlen - d;
                if (best < dend) best += match_run(s, best, d + best, dend);
                else best = dend;
            }
            LAT(L, e, d) = best;
            if (best != plen) continue;

            /* :357-413 -- can e plain mismatches explain it? */
            int straight = 0;
            for (int i = 0; i < end; i++) straight += pat_at(s, i) != txt_at(s, i);
            straight += plen - end;
            if (straight == e) {
                if (use_m) {
                    if (!cigar_put(&o, plen, 'M')) return -2;
                } else {
                    int start = 0;
                    int matching = pat_at(s, 0) == txt_at(s, 0);
                    for (int i = 0; i < end; i++) {
                        int m = pat_at(s, i) == txt_at(s, i);
                        if (m != matching) {
                            if (!cigar_put(&o, i - start, matching ? '=' : 'X')) return -2;
                            matching = m;
                            start = i;
                        }
                    }
                    if (plen > start) {
                        if (!matching) {
                            if (!cigar_put(&o, plen - start, 'X')) return -2;
                        } else {
                            if (!cigar_put(&o, end - start, '=')) return -2;
                            if (plen > end && !cigar_put(&o, plen - end, 'X')) return -2;
                        }
                    }
                }
                *(o.buf - (o.left == 0 ? 1 : 0)) = 0;
                return e;
            }
            /* backtrace, :441-531 */
            char bt_act[MAXK + 1];
            int bt_matched[MAXK + 1];
            int cur_d = d;
            for (int ce = e; ce >= 1; ce--) {
                char a = LAT(A, ce, cur_d);
                bt_act[ce] = a;
                if (a == 'I') {
                    bt_matched[ce] = LAT(L, ce, cur_d) - LAT(L, ce - 1, cur_d + 1) - 1;
                    cur_d++;
                } else if (a == 'D') {
                    bt_matched[ce] = LAT(L, ce, cur_d) - LAT(L, ce - 1, cur_d - 1);
                    cur_d--;
                } else {
                    bt_matched[ce] = LAT(L, ce, cur_d) - LAT(L, ce - 1, cur_d) - 1;
                }
            }
            int acc_m = 0;
            if (use_m) {
                acc_m = LAT(L, 0, 0);
            } else if (LAT(L, 0, 0) > 0) {
                if (!cigar_put(&o, LAT(L, 0, 0), '=')) return -2;
            }
            int ce = 1;
            while (ce <= e) {
                char a = bt_act[ce];
                int count = 1;
                while (ce + 1 <= e && bt_matched[ce] == 0 && bt_act[ce + 1] == a) { count++; ce++; }
                if (use_m) {
                    if (a == 'X') {
                        acc_m += count;
                    } else {
                        if (acc_m != 0) {
                            if (!cigar_put(&o, acc_m, 'M')) return -2;
                            acc_m = 0;
                        }
                        if (!cigar_put(&o, count, a)) return -2;
                    }
                } else {
                    if (!cigar_put(&o, count, a)) return -2;
                }
                if (bt_matched[ce] > 0) {
                    if (use_m) acc_m += bt_matched[ce];
                    else if (!cigar_put(&o, bt_matched[ce], '=')) return -2;
                }
                ce++;
            }
            if (use_m && acc_m != 0 && !cigar_put(&o, acc_m, 'M')) return -2;
            *(o.buf - (o.left == 0 ? 1 : 0)) = 0;
            return e;
        }
    }
    *(o.buf - (o.left == 0 ? 1 : 0)) = 0;
    return -1;
}

int oracle_lv_batch(int text_direction, uint32_t n, const uint32_t *text_offsets, const uint8_t *texts,
                    const uint32_t *pattern_offsets, const uint8_t *patterns, const uint8_t *quals, const int32_t *k,
                    int32_t *score, double *match_probability, int32_t *net_indel)
{
    init_tables();
    for (uint32_t i = 0; i < n; i++) {
        lv_strings s;
        s.plen = pattern_offsets[i + 1] - pattern_offsets[i];
        s.p = patterns + pattern_offsets[i];
        s.tlen = text_offsets[i + 1] - text_offsets[i];
        s.dir = text_direction;
        s.t = text_direction > 0 ? texts + text_offsets[i] : texts + text_offsets[i + 1];
        s.t_lo = 0;
        s.t_hi = s.tlen;
        double prob = 0;
        int indel = 0;
        score[i] = lv_score(&s, quals ? quals + pattern_offsets[i] : NULL, k[i], quals ? &prob : NULL, &indel);
        if (match_probability) match_probability[i] = prob;
        if (net_indel) net_indel[i] = indel;
    }
    return 0;
}

int oracle_lv_cigar_batch(uint32_t n, const uint32_t *text_offsets, const uint8_t *texts,
                          const uint32_t *pattern_offsets, const uint8_t *patterns, const int32_t *k, int use_m,
                          char *cigars, uint32_t cigar_stride, int32_t *edit_distance)
{
    for (uint32_t i = 0; i < n; i++) {
        lv_strings s;
        s.plen = pattern_offsets[i + 1] - pattern_offsets[i];
        s.p = patterns + pattern_offsets[i];
        s.tlen = text_offsets[i + 1] - text_offsets[i];
        s.dir = 1;
        s.t = texts + text_offsets[i];
        s.t_lo = 0;
        s.t_hi = s.tlen;
        char *out = cigars + (size_t)i * cigar_stride;
        memset(out, 0, cigar_stride);
        edit_distance[i] = lv_cigar(&s, k[i], out, (int)cigar_stride, use_m);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* Index + genome in host memory (file formats: GenomeIndex.cpp:646-710, HashTable.cpp:181-215,        */
/* Genome.cpp:126-158)                                                                                */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { uint32_t key, v1, v2; } ht_entry; /* HashTable.h:119-123 */

typedef struct {
    uint32_t seed_len, n_tables, overflow_words, padding;
    uint64_t *table_size;
    ht_entry **table;
    uint32_t *overflow;
    uint32_t n_bases, n_pieces;
    uint32_t *piece_begin;
    uint8_t *bases_alloc; /* 100 'n' + bases + 100 'n' (Genome.h:175, Genome.cpp:33-41) */
    uint8_t *bases;
} oracle_index;

#define GENOME_PAD 100

void oracle_index_close(void *h)
{
    oracle_index *x = (oracle_index *)h;
    if (!x) return;
    if (x->table) for (uint32_t i = 0; i < x->n_tables; i++) free(x->table[i]);
    free(x->table); free(x->table_size); free(x->overflow); free(x->piece_begin); free(x->bases_alloc); free(x);
}

void *oracle_index_load(const char *dir)
{
    init_tables();
    char path[4096];
    oracle_index *x = (oracle_index *)calloc(1, sizeof(*x));
    snprintf(path, sizeof(path), "%s/GenomeIndex", dir);
    FILE *f = fopen(path, "r");
    if (!f) { free(x); return NULL; }
    unsigned major, minor;
    if (fscanf(f, "%u %u %u %u %u %u", &major, &minor, &x->n_tables, &x->overflow_words, &x->seed_len, &x->padding) != 6) {
        fclose(f); free(x); return NULL;
    }
    fclose(f);
    snprintf(path, sizeof(path), "%s/OverflowTable", dir);
    f = fopen(path, "rb");
    if (!f) { free(x); return NULL; }
    x->overflow = (uint32_t *)malloc((size_t)x->overflow_words * 4 + 4);
    if (fread(x->overflow, 4, x->overflow_words, f) != x->overflow_words) { fclose(f); oracle_index_close(x); return NULL; }
    fclose(f);
    snprintf(path, sizeof(path), "%s/GenomeIndexHash", dir);
    f = fopen(path, "rb");
    if (!f) { oracle_index_close(x); return NULL; }
    x->table = (ht_entry **)calloc(x->n_tables, sizeof(ht_entry *));
    x->table_size = (uint64_t *)calloc(x->n_tables, sizeof(uint64_t));
    for (uint32_t i = 0; i < x->n_tables; i++) {
        uint32_t magic;
        uint64_t size, used;
        if (fread(&magic, 4, 1, f) != 1 || fread(&size, 8, 1, f) != 1 || fread(&used, 8, 1, f) != 1 || magic != 0xb111b010u) {
            fclose(f); oracle_index_close(x); return NULL;
        }
        x->table_size[i] = size;
        x->table[i] = (ht_entry *)malloc(size * sizeof(ht_entry));
        if (fread(x->table[i], sizeof(ht_entry), size, f) != size) { fclose(f); oracle_index_close(x); return NULL; }
    }
    fclose(f);
    snprintf(path, sizeof(path), "%s/Genome", dir);
    f = fopen(path, "rb");
    if (!f) { oracle_index_close(x); return NULL; }
    char line[1024];
    if (!fgets(line, sizeof(line), f) || sscanf(line, "%u %u", &x->n_bases, &x->n_pieces) != 2) { fclose(f); oracle_index_close(x); return NULL; }
    x->piece_begin = (uint32_t *)calloc(x->n_pieces ? x->n_pieces : 1, 4);
    for (uint32_t i = 0; i < x->n_pieces; i++) {
        if (!fgets(line, sizeof(line), f)) { fclose(f); oracle_index_close(x); return NULL; }
        x->piece_begin[i] = (uint32_t)atoi(line);
    }
    x->bases_alloc = (uint8_t *)malloc((size_t)x->n_bases + 2 * GENOME_PAD);
    memset(x->bases_alloc, 'n', (size_t)x->n_bases + 2 * GENOME_PAD);
    x->bases = x->bases_alloc + GENOME_PAD;
    if (fread(x->bases, 1, x->n_bases, f) != x->n_bases) { fclose(f); oracle_index_close(x); return NULL; }
    fclose(f);
    return x;
}

int oracle_index_info(void *h, snapb200_index_info *info)
{
    oracle_index *x = (oracle_index *)h;
    memset(info, 0, sizeof(*info));
    info->n_bases = x->n_bases; info->n_pieces = x->n_pieces; info->seed_len = x->seed_len;
    info->n_hash_tables = x->n_tables; info->overflow_table_size = x->overflow_words;
    info->chromosome_padding = x->padding;
    for (uint32_t i = 0; i < x->n_tables; i++) info->hash_table_entries += x->table_size[i];
    info->device = -1;
    return 0;
}

/* Seed::Seed + DoesTextRepresentASeed: Seed.h:38-51, Seed.cpp:29-42, Tables.cpp:36-42 (A=0 G=1 C=2 T=3) */
static inline int base2(uint8_t c) { return c == 'A' ? 0 : c == 'G' ? 1 : c == 'C' ? 2 : c == 'T' ? 3 : -1; }

static int pack_seed(const uint8_t *text, uint32_t len, uint64_t *fwd, uint64_t *rc)
{
    uint64_t f = 0, r = 0;
    for (uint32_t i = 0; i < len; i++) {
        int v = base2(text[i]);
        if (v < 0) return 0;
        f |= (uint64_t)v << ((len - i - 1) * 2);
        r |= (uint64_t)(v ^ 3) << (i * 2);
    }
    *fwd = f; *rc = r;
    return 1;
}

static inline uint32_t ht_hash(uint32_t key) /* HashTable.h:60-72 */
{
    key ^= key >> 16; key *= 0x85ebca6bu; key ^= key >> 13; key *= 0xc2b2ae35u; key ^= key >> 16;
    return key;
}

/* SNAPHashTable::Lookup, HashTable.h:74-105.  *probes gets the number of slots examined. */
static const ht_entry *ht_lookup(const ht_entry *t, uint64_t size, uint32_t key, uint32_t *probes)
{
    uint64_t idx = ht_hash(key) % size;
    *probes = 1;
    if (t[idx].key == key && t[idx].v1 != INVALID_LOC) return &t[idx];
    uint64_t n = 0;
    const ht_entry *e;
    do {
        n++;
        if (n > size + 5) return NULL;
        if (n < 5) idx = (idx + n * n) % size; else idx = (idx + 1) % size;
        e = &t[idx];
        (*probes)++;
    } while (e->key != key && e->v1 != INVALID_LOC);
    return e->v1 == INVALID_LOC ? NULL : e;
}

typedef struct { uint32_t n; const uint32_t *hits; } hit_list;

/* GenomeIndex::fillInLookedUpResults, GenomeIndex.cpp:1013-1086 (full-range form) */
static hit_list resolve_hits(const oracle_index *x, const uint32_t *sub)
{
    hit_list r = {0, NULL};
    if (*sub < x->n_bases) { r.n = 1; r.hits = sub; }
    else if (*sub == 0xfffffffeu) { r.n = 0; }
    else {
        uint32_t off = *sub - x->n_bases;
        r.n = x->overflow[off];
        r.hits = &x->overflow[off + 1];
    }
    return r;
}

/* GenomeIndex::lookupSeed, GenomeIndex.cpp:971-1011 */
static void lookup_seed(const oracle_index *x, uint64_t fwd, uint64_t rc, hit_list out[2], uint32_t *probes)
{
    int swapped = (int64_t)fwd > (int64_t)rc;
    uint64_t s = swapped ? rc : fwd;
    uint32_t hi = (uint32_t)(s >> 32), lo = (uint32_t)s;
    out[0].n = out[1].n = 0; out[0].hits = out[1].hits = NULL;
    uint32_t pr = 0;
    const ht_entry *e = ht_lookup(x->table[hi], x->table_size[hi], lo, &pr);
    if (probes) *probes = pr;
    if (!e) return;
    out[0] = resolve_hits(x, swapped ? &e->v2 : &e->v1);
    if (fwd == rc) out[1] = out[0];
    else out[1] = resolve_hits(x, swapped ? &e->v1 : &e->v2);
}

int oracle_lookup_seed_batch(void *h, uint32_t n, const uint8_t *seeds, uint32_t max_out, uint32_t *n_hits, uint32_t *hits)
{
    oracle_index *x = (oracle_index *)h;
    for (uint32_t i = 0; i < n; i++) {
        uint64_t f, r;
        hit_list hl[2] = {{0, NULL}, {0, NULL}};
        if (pack_seed(seeds + (size_t)i * x->seed_len, x->seed_len, &f, &r)) lookup_seed(x, f, r, hl, NULL);
        for (int d = 0; d < 2; d++) {
            n_hits[i * 2 + d] = hl[d].n;
            for (uint32_t j = 0; j < hl[d].n && j < max_out; j++) hits[((size_t)i * 2 + d) * max_out + j] = hl[d].hits[j];
        }
    }
    return 0;
}

/* GetWrappedNextSeedToTest, SeedSequencer.h:28-287: the wrap order, as data (row = seedLen-16). */
static const uint8_t WRAP_ORDER[10][25] = {
    {0, 8, 4, 12, 2, 6, 10, 14, 1, 3, 5, 7, 9, 11, 13, 15},
    {0, 8, 4, 12, 2, 6, 10, 14, 1, 3, 5, 7, 9, 11, 13, 15, 16},
    {0, 9, 4, 13, 2, 6, 11, 15, 1, 3, 5, 7, 8, 10, 12, 14, 16, 17},
    {0, 10, 4, 14, 2, 6, 8, 12, 16, 18, 1, 3, 5, 7, 9, 11, 13, 15, 17},
    {0, 10, 5, 15, 2, 7, 12, 17, 3, 9, 11, 13, 19, 1, 4, 6, 8, 14, 18, 16},
    {0, 11, 6, 16, 3, 9, 13, 17, 18, 2, 5, 8, 15, 20, 1, 4, 7, 10, 12, 14, 19},
    {0, 11, 6, 16, 3, 9, 14, 19, 2, 7, 12, 17, 20, 4, 1, 10, 13, 15, 18, 21, 5, 8},
    {0, 12, 6, 17, 3, 9, 20, 14, 1, 4, 7, 10, 15, 18, 21, 4, 2, 5, 11, 16, 19, 22, 8}, /* sic: 4 twice */
    {0, 12, 6, 18, 3, 15, 21, 9, 1, 13, 19, 7, 16, 4, 22, 10, 2, 14, 20, 5, 17, 8, 23, 11},
    {0, 13, 6, 19, 3, 16, 22, 9, 11, 1, 14, 7, 20, 4, 17, 23, 2, 15, 5, 21, 8, 24, 10, 18, 12},
};
static inline uint32_t wrapped_seed(uint32_t seed_len, uint32_t wrap) { return WRAP_ORDER[seed_len - 16][wrap]; }

/* Genome::getSubstring, Genome.h:78-148, restricted to lengthNeeded <= chromosomePadding (the only case the
 * aligners produce for reads <= 469 bases with the default 500 padding).  Returns an offset or -1. */
static int genome_substring_ok(const oracle_index *x, uint32_t offset, uint32_t len)
{
    if ((uint64_t)offset > x->n_bases || (uint64_t)offset + len > (uint64_t)x->n_bases + GENOME_PAD) return 0;
    return 1;
}

static uint32_t piece_begin_after(const oracle_index *x, uint32_t loc)
{ /* beginningOffset of the first piece that starts after loc (Genome::getNextPieceAfterLocation) */
    for (uint32_t i = 0; i < x->n_pieces; i++) if (x->piece_begin[i] > loc) return x->piece_begin[i];
    return x->n_bases;
}
static uint32_t piece_begin_at(const oracle_index *x, uint32_t loc)
{ /* Genome::getPieceAtLocation(loc)->beginningOffset */
    uint32_t b = 0;
    for (uint32_t i = 0; i < x->n_pieces; i++) if (x->piece_begin[i] <= loc) b = x->piece_begin[i];
    return b;
}

/* a read in both orientations, as the aligners prepare it (BaseAligner.cpp:638-661,
 * IntersectingPairedEndAligner.cpp:193-241) */
typedef struct {
    uint32_t len;
    uint8_t data[2][SNAPB200_MAX_READ_LENGTH + 8];     /* [FORWARD] read, [RC] reverse complement   */
    uint8_t qual[2][SNAPB200_MAX_READ_LENGTH + 8];     /* [RC] = reversed quality                   */
    uint8_t reversed[2][SNAPB200_MAX_READ_LENGTH + 8]; /* data[dir] reversed, for the backward LV   */
    uint32_t n_count;
} read_views;

static inline uint8_t rc_base(uint8_t c)
{ /* rcTranslationTable, BaseAligner.cpp:148-152; other bytes are unspecified in the reference: map to 0 */
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; case 'N': return 'N'; }
    return 0;
}

static void make_views(read_views *v, const uint8_t *bases, const uint8_t *quals, uint32_t len)
{
    v->len = len;
    v->n_count = 0;
    for (uint32_t i = 0; i < len; i++) {
        uint8_t b = bases[i], c = rc_base(b);
        v->data[0][i] = b;
        v->qual[0][i] = quals[i];
        v->data[1][len - 1 - i] = c;
        v->qual[1][len - 1 - i] = quals[i];
        v->reversed[0][len - 1 - i] = b;
        v->reversed[1][i] = c;
        v->n_count += b == 'N';
    }
}

/* The scoring step shared by BaseAligner::score (BaseAligner.cpp:1158-1242) and
 * IntersectingPairedEndAligner::scoreLocation (:755-841).  `single_variant` selects which of the two
 * (slightly different) end-of-contig rules applies.  Returns score or -1. */
static int score_location(const oracle_index *x, const read_views *v, int dir, uint32_t loc, uint32_t seed_offset,
                          int score_limit, int single_variant, double seed_prob, double *match_prob, int *loc_offset)
{
    uint32_t rlen = v->len;
    uint32_t glen = rlen + MAXK;
    int have = genome_substring_ok(x, loc, glen);
    if (!have) {
        uint32_t end_off;
        if ((uint64_t)loc + rlen + MAXK >= x->n_bases) end_off = x->n_bases;
        else end_off = single_variant ? piece_begin_after(x, loc) : piece_begin_at(x, loc + rlen + MAXK);
        glen = end_off - loc - 1;
        if (glen >= rlen - MAXK) have = genome_substring_ok(x, loc, glen);
    }
    *match_prob = 0;
    *loc_offset = 0;
    if (!have) return -1;
    const uint8_t *data = x->bases + loc;
    int seed_len = (int)x->seed_len;
    int tail = (int)seed_offset + seed_len;
    lv_strings s;
    double p1, p2;
    /* forward: read tail vs genome after the seed */
    s.p = v->data[dir] + tail; s.plen = (int)rlen - tail;
    s.t = data + tail; s.tlen = (int)glen - tail; s.dir = 1;
    s.t_lo = -(int)(loc + tail) - GENOME_PAD; s.t_hi = (int)((int64_t)x->n_bases + GENOME_PAD - loc - tail);
    int s1 = lv_score(&s, v->qual[dir] + tail, score_limit, &p1, NULL);
    if (s1 == -1) return -1;
    /* backward: reversed read head vs genome before the seed */
    s.p = v->reversed[dir] + rlen - seed_offset; s.plen = (int)seed_offset;
    s.t = data + seed_offset; s.tlen = (int)seed_offset + MAXK; s.dir = -1;
    s.t_lo = -(int)((int64_t)x->n_bases + GENOME_PAD - loc - seed_offset); s.t_hi = (int)(loc + seed_offset) + GENOME_PAD;
    int s2 = lv_score(&s, v->qual[1 - dir] + rlen - seed_offset, score_limit - s1, &p2, loc_offset);
    if (s2 == -1) { *loc_offset = 0; return -1; }
    *match_prob = p1 * p2 * seed_prob;
    return s1 + s2;
}

/* ------------------------------------------------------------------------------------------------ */
/* BaseAligner (single end)                                                                           */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    int w_next, w_prev;  /* weight list links (indices; negative = list head -(w+1)) */
    int h_next;          /* hash chain */
    uint64_t used, scored;
    uint32_t base, weight, lowest_possible, best_score, best_loc;
    uint8_t dir, all_scored;
    double best_prob;
    uint16_t seed_offset[BUCKET];
} element;

typedef struct {
    const oracle_index *x;
    snapb200_single_params p;
    uint32_t n_lists, pool_size, table_size;
    element *pool;
    uint32_t n_used;
    int *anchor[2];       /* candidate hash table: head element index or -1 */
    int *list_head;       /* per weight: first element, -1 empty */
    int *list_tail;
    uint32_t highest_list;
    double seed_prob;
    /* multi-hit capture, BaseAligner.h:149-152 (flat layout so row overflow lands where it does there) */
    uint32_t hit_count[MAXK];
    uint32_t hit_loc[MAXK * 512];
    uint8_t hit_rc[MAXK * 512];
    /* per-read state */
    uint32_t lowest_unseen[2], most_seeds[2], n_applied[2];
    uint32_t best_score, best_loc, second_best, second_loc, score_limit, popular_skipped;
    double p_all, p_best;
    uint32_t n_lookups, n_scored;
} base_aligner;

static uint32_t seeds_to_use(uint32_t num_seeds, double coverage, uint32_t read_size, uint32_t seed_len)
{ /* BaseAligner.cpp:120-125, 563-568 */
    if (num_seeds != 0) return num_seeds;
    return (uint32_t)(int)(coverage * read_size / seed_len);
}

static base_aligner *ba_create(const oracle_index *x, const snapb200_single_params *p)
{
    base_aligner *a = (base_aligner *)calloc(1, sizeof(*a));
    a->x = x; a->p = *p;
    uint32_t max_seeds = seeds_to_use(p->num_seeds, p->seed_coverage, p->max_read_size, x->seed_len);
    a->n_lists = max_seeds + 1;                     /* BaseAligner.cpp:127 */
    a->table_size = (p->max_hits * max_seeds * 3) / 2; /* :129 */
    a->pool_size = p->max_hits * max_seeds * 2;      /* :130 */
    if (a->table_size == 0) a->table_size = 1;
    a->pool = (element *)malloc(sizeof(element) * (a->pool_size ? a->pool_size : 1));
    for (int d = 0; d < 2; d++) a->anchor[d] = (int *)malloc(sizeof(int) * a->table_size);
    a->list_head = (int *)malloc(sizeof(int) * a->n_lists);
    a->list_tail = (int *)malloc(sizeof(int) * a->n_lists);
    a->seed_prob = powi_ref(1 - 0.001, (int)x->seed_len); /* BaseAligner.cpp:1227 */
    return a;
}
static void ba_destroy(base_aligner *a)
{
    free(a->pool); free(a->anchor[0]); free(a->anchor[1]); free(a->list_head); free(a->list_tail); free(a);
}

/* weight lists are FIFO doubly linked lists (BaseAligner.cpp:1549-1552, 1714-1726) */
static void list_unlink(base_aligner *a, int e)
{
    element *el = &a->pool[e];
    if (el->w_next == e && el->w_prev == e) return; /* self-linked: not on any list (:1394) */
    uint32_t w = el->weight;
    if (el->w_prev >= 0) a->pool[el->w_prev].w_next = el->w_next; else a->list_head[w] = el->w_next;
    if (el->w_next >= 0) a->pool[el->w_next].w_prev = el->w_prev; else a->list_tail[w] = el->w_prev;
    el->w_next = el->w_prev = e;
}
static void list_append(base_aligner *a, int e, uint32_t w)
{
    element *el = &a->pool[e];
    el->w_next = -1;
    el->w_prev = a->list_tail[w];
    if (a->list_tail[w] >= 0) a->pool[a->list_tail[w]].w_next = e; else a->list_head[w] = e;
    a->list_tail[w] = e;
}

static int ba_find_element(base_aligner *a, uint32_t loc, int dir)
{ /* findElement, BaseAligner.cpp:1415-1442 */
    uint32_t base = loc - loc % BUCKET;
    int e = a->anchor[dir][(base * 131u) % a->table_size];
    while (e >= 0 && a->pool[e].base != base) e = a->pool[e].h_next;
    return e;
}

static void ba_increment_weight(base_aligner *a, int e)
{ /* incrementWeight, BaseAligner.cpp:1689-1727 */
    element *el = &a->pool[e];
    if (el->all_scored) return;
    if (el->weight >= a->n_lists - 1) return;
    int on_list = !(el->w_next == e && el->w_prev == e);
    if (on_list) list_unlink(a, e);
    el->weight++;
    if (el->weight > a->highest_list) a->highest_list = el->weight;
    list_append(a, e, el->weight);
}

static void ba_new_candidate(base_aligner *a, uint32_t loc, int dir, uint32_t lowest, uint32_t seed_offset)
{ /* allocateNewCandidate, BaseAligner.cpp:1485-1568 */
    uint32_t low = loc % BUCKET, base = loc - low;
    int e = (int)a->n_used++;
    element *el = &a->pool[e];
    el->used = (uint64_t)1 << low;
    el->scored = 0;
    el->lowest_possible = lowest;
    el->dir = (uint8_t)dir;
    el->weight = 1;
    el->base = base;
    el->best_score = UNUSED_SCORE;
    el->all_scored = 0;
    el->best_prob = 0;
    el->best_loc = 0;
    list_append(a, e, 1);
    el->seed_offset[low] = (uint16_t)seed_offset;
    if (a->highest_list < 1) a->highest_list = 1;
    uint32_t slot = (base * 131u) % a->table_size;
    el->h_next = a->anchor[dir][slot];
    a->anchor[dir][slot] = e;
}

typedef struct {
    uint32_t *location;
    int *direction;
    int *final_score;
    int *mapq;
    int status;
} ba_out;

/* BaseAligner::score, BaseAligner.cpp:977-1399.  Returns 1 when a final answer was produced. */
static int ba_score(base_aligner *a, int force, const read_views *v, ba_out *o)
{
    const uint32_t max_k = a->p.max_k, extra = a->p.extra_search_depth;
    for (int d = 0; d < 2; d++) {
        if (a->most_seeds[d] != 0) {
            uint32_t q = a->n_applied[d] / a->most_seeds[d];
            if (q > a->lowest_unseen[d]) a->lowest_unseen[d] = q;
        }
    }
    uint32_t list = a->highest_list;
    do {
        while (list > 0 && a->list_head[list] < 0) { list--; a->highest_list = list; }
        uint32_t lo = a->lowest_unseen[0] < a->lowest_unseen[1] ? a->lowest_unseen[0] : a->lowest_unseen[1];
        if (lo > a->score_limit || force) {
            if (list == 0) {
                *o->final_score = (int)a->best_score;
                if (a->best_score <= max_k) {
                    *o->location = a->best_loc;
                    *o->mapq = compute_mapq(a->p_all, a->p_best, (int)a->best_score, (int)a->popular_skipped);
                    o->status = *o->mapq >= 10 ? SNAPB200_SINGLE_HIT : SNAPB200_MULTIPLE_HITS;
                } else {
                    o->status = (a->n_applied[0] == 0 && a->n_applied[1] == 0) ? SNAPB200_MULTIPLE_HITS : SNAPB200_NOT_FOUND;
                    *o->mapq = 0;
                }
                return 1;
            }
            force = 1;
        } else if (list == 0) {
            return 0;
        }
        int ei = a->list_head[list];
        element *el = &a->pool[ei];
        if (el->lowest_possible <= a->score_limit) {
            uint64_t mask = el->used; /* snapshot, :1132 */
            while (mask) {
                unsigned idx = (unsigned)__builtin_ctzll(mask);
                uint64_t bit = (uint64_t)1 << idx;
                mask &= ~bit;
                if (el->scored & bit) continue;
                int any_nearby = el->scored != 0;
                el->scored |= bit;
                uint32_t loc = el->base + idx;
                uint32_t elem_loc = loc;
                double prob = 0;
                int loc_off = 0;
                int sc = score_location(a->x, &v[0], el->dir, loc, el->seed_offset[idx], (int)a->score_limit, 1, a->seed_prob,
                                        &prob, &loc_off);
                uint32_t score = (uint32_t)sc; /* -1 -> 0xffffffff like the reference's unsigned */
                if (sc != -1) loc += (uint32_t)loc_off;
                if (a->p.max_hits_to_get > 0 && sc != -1 && a->hit_count[score] < a->p.max_hits_to_get) { /* :1255-1261 */
                    uint32_t flat = score * 512 + a->hit_count[score];
                    if (flat < MAXK * 512) { a->hit_loc[flat] = loc; a->hit_rc[flat] = el->dir; }
                    a->hit_count[score]++;
                }
                a->n_scored++;
                if (any_nearby) { /* :1272-1278 */
                    if (el->best_score < score || (el->best_score == score && prob <= el->best_prob)) continue;
                }
                el->best_loc = loc;
                int near = -1;
                if (sc != -1) { /* :1298-1305 */
                    uint32_t half = BUCKET / 2;
                    uint32_t near_loc = elem_loc + (2 * (elem_loc % BUCKET / half) - 1) * half;
                    near = ba_find_element(a, near_loc, el->dir);
                }
                if (near >= 0 && a->pool[near].scored != 0) {
                    element *ne = &a->pool[near];
                    if (!((ne->base > el->base && loc - ne->best_loc <= BUCKET) ||
                          (ne->base < el->base && ne->best_loc <= BUCKET))) { /* :1311-1312, quirk kept */
                        near = -1;
                    }
                    if (near >= 0) {
                        if (ne->best_score < score || (ne->best_score == score && ne->best_prob >= prob)) continue;
                        any_nearby = 1;
                        double t = a->p_all - ne->best_prob;
                        a->p_all = t > 0.0 ? t : 0.0;
                        ne->best_prob = 0;
                    }
                }
                {
                    double t = a->p_all - el->best_prob;
                    a->p_all = t > 0.0 ? t : 0.0;
                }
                a->p_all += prob;
                el->best_prob = prob;
                el->best_score = score;
                if (a->best_score > score || (a->best_score == score && prob > a->p_best)) {
                    /* second-best bookkeeping (:1345-1352) does not reach any output; kept for fidelity */
                    if ((a->second_best == UNUSED_SCORE || !(a->second_loc + BUCKET > loc && a->second_loc < loc + BUCKET)) &&
                        (a->best_score == UNUSED_SCORE || !(a->best_loc + BUCKET > loc && a->best_loc < loc + BUCKET)) &&
                        (!any_nearby || (a->best_loc / BUCKET != loc / BUCKET && a->second_loc / BUCKET != loc / BUCKET))) {
                        a->second_best = a->best_score;
                        a->second_loc = a->best_loc;
                    }
                    a->best_score = score;
                    a->p_best = prob;
                    a->best_loc = loc;
                    *o->location = loc;
                    *o->final_score = (int)score;
                    *o->direction = el->dir;
                } else if (a->second_best > score) {
                    a->second_best = score;
                    a->second_loc = loc;
                }
                if (a->p.stop_on_first_hit && a->best_score <= max_k) { /* :1373-1381 */
                    o->status = SNAPB200_MULTIPLE_HITS;
                    *o->mapq = 0;
                    return 1;
                }
                a->score_limit = (a->best_score < max_k ? a->best_score : max_k) + extra; /* :1384 */
            }
        }
        el->all_scored = 1; /* :1391-1394 */
        list_unlink(a, ei);
    } while (force);
    return 0;
}

static void ba_fill_hits(base_aligner *a, int32_t *found, uint32_t *locs, uint8_t *rcs, int32_t *scores)
{ /* fillHitsFound, BaseAligner.cpp:940-975 */
    uint32_t want = a->p.max_hits_to_get;
    if (want == 0) return;
    *found = 0;
    int first = 0;
    while (first < MAXK && a->hit_count[first] == 0) first++;
    int last = first + 4 < MAXK ? first + 4 : MAXK;
    for (int dist = first; dist < last; dist++) {
        for (uint32_t i = 0; i < a->hit_count[dist]; i++) {
            uint32_t flat = (uint32_t)dist * 512 + i;
            locs[*found] = flat < MAXK * 512 ? a->hit_loc[flat] : 0;
            rcs[*found] = flat < MAXK * 512 ? a->hit_rc[flat] : 0;
            scores[*found] = dist;
            *found += 1;
            if ((uint32_t)*found == want) return;
        }
    }
}

/* BaseAligner::AlignRead, BaseAligner.cpp:510-938 (searchRadius == 0) */
static void ba_align(base_aligner *a, const uint8_t *bases, const uint8_t *quals, uint32_t len, snapb200_single_result *r,
                     int32_t *found, uint32_t *hlocs, uint8_t *hrcs, int32_t *hscores)
{
    const oracle_index *x = a->x;
    const uint32_t seed_len = x->seed_len;
    uint32_t location = INVALID_LOC;
    int direction = SNAPB200_FORWARD, final_score = UNUSED_SCORE, mapq = 0;
    ba_out o = {&location, &direction, &final_score, &mapq, SNAPB200_NOT_FOUND};
    memset(r, 0, sizeof(*r));
    a->p_all = a->p_best = 0;
    a->popular_skipped = 0;
    a->n_lookups = a->n_scored = 0;
    if (a->p.max_hits_to_get > 0) { memset(a->hit_count, 0, sizeof(a->hit_count)); *found = 0; }
    uint32_t max_seeds = seeds_to_use(a->p.num_seeds, a->p.seed_coverage, len, seed_len);
    static __thread read_views v;
    int done = 0;
    if (len < seed_len) { done = 1; }
    if (!done) {
        make_views(&v, bases, quals, len);
        if (v.n_count > a->p.max_k) done = 1;
    }
    if (!done) {
        /* clearCandidates, :1679-1687 */
        a->n_used = 0;
        a->highest_list = 0;
        for (uint32_t i = 0; i < a->n_lists; i++) a->list_head[i] = a->list_tail[i] = -1;
        for (int d = 0; d < 2; d++) for (uint32_t i = 0; i < a->table_size; i++) a->anchor[d][i] = -1;
        uint8_t used[SNAPB200_MAX_READ_LENGTH + 8];
        memset(used, 0, sizeof(used));
        uint32_t n_possible = len - seed_len + 1, next = 0, wrap = 0;
        a->lowest_unseen[0] = a->lowest_unseen[1] = 0;
        a->most_seeds[0] = a->most_seeds[1] = 1;
        a->best_score = a->second_best = UNUSED_SCORE;
        a->best_loc = a->second_loc = 0;
        a->n_applied[0] = a->n_applied[1] = 0;
        a->score_limit = a->p.max_k + a->p.extra_search_depth;
        int finished = 0;
        while (a->n_applied[0] + a->n_applied[1] < max_seeds) {
            if (next >= n_possible) { /* :690-723 */
                wrap++;
                if (wrap >= seed_len) {
                    ba_score(a, 1, &v, &o);
                    finished = 2; /* NB: this exit does not call fillHitsFound (:705-718) */
                    break;
                }
                next = wrapped_seed(seed_len, wrap);
                a->most_seeds[0] = a->most_seeds[1] = wrap + 1;
            }
            while (next < n_possible && used[next]) next++;
            if (next >= n_possible) continue;
            used[next] = 1;
            uint64_t sf, sr;
            if (!pack_seed(v.data[0] + next, seed_len, &sf, &sr)) continue; /* :742-744 */
            hit_list hl[2];
            lookup_seed(x, sf, sr, hl, NULL);
            a->n_lookups++;
            int applied = 0;
            for (int dir = 0; dir < 2; dir++) {
                if (hl[dir].n > a->p.max_hits && !a->p.explore_popular_seeds) {
                    a->popular_skipped++;
                    continue;
                }
                uint32_t offset = dir == 0 ? next : len - seed_len - next;
                uint32_t lim = hl[dir].n < a->p.max_hits ? hl[dir].n : a->p.max_hits;
                for (uint32_t i = 0; i < lim; i++) {
                    uint32_t hit = hl[dir].hits[i];
                    if (hit < offset) continue; /* :848-853 with the full window */
                    uint32_t loc = hit - offset;
                    int e = ba_find_element(a, loc, dir);
                    if (e >= 0) { /* findCandidate, :1445-1481 */
                        element *el = &a->pool[e];
                        uint64_t bit = (uint64_t)1 << (loc % BUCKET);
                        el->all_scored = el->all_scored && (el->used & bit) != 0;
                        el->used |= bit;
                        ba_increment_weight(a, e);
                        el->seed_offset[loc % BUCKET] = (uint16_t)offset;
                    } else if (a->lowest_unseen[dir] <= a->score_limit) {
                        ba_new_candidate(a, loc, dir, a->lowest_unseen[dir], offset);
                    }
                }
                a->n_applied[dir]++;
                applied = 1;
            }
            next += seed_len; /* :876 */
            if (applied && ba_score(a, 0, &v, &o)) { finished = 1; break; }
        }
        if (!finished) { ba_score(a, 1, &v, &o); finished = 1; }
        if (finished == 1) ba_fill_hits(a, found, hlocs, hrcs, hscores);
    }
    r->location = location;
    r->direction = (uint8_t)direction;
    r->score = final_score;
    r->mapq = mapq;
    r->status = (uint8_t)o.status;
    r->popular_seeds_skipped = (uint16_t)a->popular_skipped;
    r->n_lookups = a->n_lookups;
    r->n_scored = a->n_scored;
    r->p_all = a->p_all;
    r->p_best = a->p_best;
}

int oracle_single_multihit_batch(void *h, const snapb200_single_params *p, const snapb200_read_batch *reads,
                                 snapb200_single_result *res, int32_t *hit_counts, uint32_t *hit_locations,
                                 uint8_t *hit_rcs, int32_t *hit_scores)
{
    init_tables();
    base_aligner *a = ba_create((const oracle_index *)h, p);
    uint32_t mh = p->max_hits_to_get;
    for (uint32_t i = 0; i < reads->n; i++) {
        uint32_t off = reads->offsets[i], len = reads->offsets[i + 1] - off;
        ba_align(a, reads->bases + off, reads->quals + off, len, &res[i], mh ? &hit_counts[i] : NULL,
                 mh ? hit_locations + (size_t)i * mh : NULL, mh ? hit_rcs + (size_t)i * mh : NULL,
                 mh ? hit_scores + (size_t)i * mh : NULL);
    }
    ba_destroy(a);
    return 0;
}

int oracle_single_batch(void *h, const snapb200_single_params *p, const snapb200_read_batch *reads,
                        snapb200_single_result *res)
{
    snapb200_single_params q = *p;
    q.max_hits_to_get = 0;
    return oracle_single_multihit_batch(h, &q, reads, res, NULL, NULL, NULL, NULL);
}

/* ------------------------------------------------------------------------------------------------ */
/* BaseAligner::CharacterizeSeeds (BaseAligner.cpp:206-508)                                           */
/* ------------------------------------------------------------------------------------------------ */
/* The reference inserts (location -> seed offset) into one std::map<unsigned, std::set<unsigned>> per
 * direction; the tuples of a map in ascending (location, seedOffset) order are its in-order traversal.
 * Segment 2*i+dir of the output holds them for read i.  Two passes over the reads when the caller wants
 * the tuples: seg_offsets is always complete. */
typedef struct { uint32_t loc; uint16_t off; } char_tuple;
static int char_tuple_cmp(const void *pa, const void *pb)
{
    const char_tuple *a = (const char_tuple *)pa, *b = (const char_tuple *)pb;
    if (a->loc != b->loc) return a->loc < b->loc ? -1 : 1;
    return (int)a->off - (int)b->off;
}
typedef struct { char_tuple *t; size_t n, cap; } char_vec;
static void char_push(char_vec *v, uint32_t loc, uint32_t off)
{
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 256; v->t = (char_tuple *)realloc(v->t, v->cap * sizeof(char_tuple)); }
    v->t[v->n].loc = loc; v->t[v->n].off = (uint16_t)off; v->n++;
}

static void characterize_read(const oracle_index *x, const snapb200_single_params *p, const uint8_t *bases, uint32_t len,
                              char_vec out[2])
{
    const uint32_t seed_len = x->seed_len;
    out[0].n = out[1].n = 0;
    if (len < seed_len) return; /* :277-282 */
    uint32_t n_count = 0;
    for (uint32_t i = 0; i < len; i++) n_count += bases[i] == 'N';
    if (n_count > p->max_k) return; /* :303-306 */
    const uint32_t max_seeds = seeds_to_use(p->num_seeds, p->seed_coverage, len, seed_len);
    uint8_t used[SNAPB200_MAX_READ_LENGTH + 8];
    memset(used, 0, sizeof(used));
    const uint32_t n_possible = len - seed_len + 1;
    uint32_t next = 0, wrap = 0, applied[2] = {0, 0};
    while (applied[0] + applied[1] < max_seeds) {
        if (next >= n_possible) { /* :330-342 */
            wrap++;
            if (wrap >= seed_len) return;
            next = wrapped_seed(seed_len, wrap);
        }
        while (next < n_possible && used[next]) next++;
        if (next >= n_possible) continue;
        used[next] = 1;
        uint64_t sf, sr;
        if (!pack_seed(bases + next, seed_len, &sf, &sr)) continue; /* :362-364 */
        hit_list hl[2];
        lookup_seed(x, sf, sr, hl, NULL);
        for (int dir = 0; dir < 2; dir++) {
            if (hl[dir].n > p->max_hits && !p->explore_popular_seeds) continue; /* :394-401 */
            const uint32_t offset = dir == 0 ? next : len - seed_len - next;
            const uint32_t lim = hl[dir].n < p->max_hits ? hl[dir].n : p->max_hits;
            for (uint32_t i = 0; i < lim; i++) {
                const uint32_t hit = hl[dir].hits[i];
                if (hit < offset) continue; /* :446-450 */
                char_push(&out[dir], hit - offset, next); /* the FORWARD offset goes into both maps, :455-478 */
            }
            applied[dir]++;
        }
        next += seed_len; /* :497 */
    }
}

int oracle_characterize_batch(void *h, const snapb200_single_params *p, const snapb200_read_batch *reads,
                              uint64_t *seg_offsets, uint32_t *locations, uint16_t *seed_offsets, uint64_t capacity)
{
    const oracle_index *x = (const oracle_index *)h;
    char_vec v[2] = {{NULL, 0, 0}, {NULL, 0, 0}};
    uint64_t pos = 0;
    int rc = 0;
    seg_offsets[0] = 0;
    for (uint32_t i = 0; i < reads->n; i++) {
        const uint32_t off = reads->offsets[i], len = reads->offsets[i + 1] - off;
        characterize_read(x, p, reads->bases + off, len, v);
        for (int dir = 0; dir < 2; dir++) {
            qsort(v[dir].t, v[dir].n, sizeof(char_tuple), char_tuple_cmp);
            if (locations) {
                if (pos + v[dir].n > capacity) { rc = SNAPB200_ERR_ARG; goto done; }
                for (size_t q = 0; q < v[dir].n; q++) { locations[pos + q] = v[dir].t[q].loc; seed_offsets[pos + q] = v[dir].t[q].off; }
            }
            pos += v[dir].n;
            seg_offsets[2 * (size_t)i + dir + 1] = pos;
        }
    }
done:
    free(v[0].t); free(v[1].t);
    return rc;
}

/* ------------------------------------------------------------------------------------------------ */
/* IntersectingPairedEndAligner + ChimericPairedEndAligner                                            */
/* ------------------------------------------------------------------------------------------------ */
#define MAX_LOOKUPS 64

typedef struct { /* HashTableLookup, IntersectingPairedEndAligner.h:101-134 */
    uint32_t seed_offset, n_hits, set, cur;
    const uint32_t *hits;
} lookup_t;

typedef struct { /* HashTableHitSet, :139-194 */
    lookup_t lk[MAX_LOOKUPS];
    uint32_t n_lookups;
    int cur_set;
    uint32_t exhausted[MAX_LOOKUPS];
    uint32_t most_recent;
    uint32_t merge_dist;
} hit_set;

static inline int is_within(uint32_t a, uint32_t b, uint32_t dist)
{ /* Util.h:538-541, unsigned wrap-around kept */
    return (a <= b && (uint32_t)(a + dist) >= b) || (a >= b && a <= (uint32_t)(b + dist));
}

static void hs_record(hit_set *s, uint32_t seed_offset, uint32_t n, const uint32_t *hits, int begins_set)
{ /* recordLookup, IntersectingPairedEndAligner.cpp:859-899 */
    if (begins_set) { s->cur_set++; s->exhausted[s->cur_set] = 0; }
    if (n == 0) { s->exhausted[s->cur_set]++; return; }
    lookup_t *l = &s->lk[s->n_lookups++];
    l->cur = 0; l->hits = hits; l->n_hits = n; l->seed_offset = seed_offset; l->set = (uint32_t)s->cur_set;
    while (l->n_hits > 0 && l->hits[l->n_hits - 1] < seed_offset) l->n_hits--;
}

static uint32_t hs_best_possible(hit_set *s)
{ /* computeBestPossibleScoreForCurrentHit, :901-929 */
    uint32_t miss[MAX_LOOKUPS];
    for (int i = 0; i <= s->cur_set; i++) miss[i] = s->exhausted[i];
    for (uint32_t i = 0; i < s->n_lookups; i++) {
        lookup_t *l = &s->lk[i];
        uint32_t target = s->most_recent + l->seed_offset;
        int close = (l->cur != l->n_hits && is_within(l->hits[l->cur], target, s->merge_dist)) ||
                    (l->cur != 0 && is_within(l->hits[l->cur - 1], target, s->merge_dist));
        if (!close) miss[l->set]++;
    }
    uint32_t best = 0;
    for (int i = 0; i <= s->cur_set; i++) if (miss[i] > best) best = miss[i];
    return best;
}

static int hs_next_le(hit_set *s, uint32_t max_loc, uint32_t *loc, uint32_t *seed_off)
{ /* getNextHitLessThanOrEqualTo, the "traditional" branch :1219-1263 */
    int any = 0;
    uint32_t best = 0;
    for (uint32_t i = 0; i < s->n_lookups; i++) {
        lookup_t *l = &s->lk[i];
        int lo = (int)l->cur, hi = (int)l->n_hits - 1;
        uint32_t want = max_loc + l->seed_offset;
        while (lo <= hi) {
            int probe = (lo + hi) / 2;
            if (l->hits[probe] <= want && (probe == 0 || l->hits[probe - 1] > want)) {
                if (l->hits[probe] - l->seed_offset > best) {
                    any = 1;
                    s->most_recent = *loc = best = l->hits[probe] - l->seed_offset;
                    *seed_off = l->seed_offset;
                }
                l->cur = (uint32_t)probe;
                break;
            }
            if (l->hits[probe] > want) lo = probe + 1; else hi = probe - 1;
        }
        if (lo > hi) l->cur = l->n_hits;
    }
    return any;
}

static int hs_first(hit_set *s, uint32_t *loc, uint32_t *seed_off)
{ /* getFirstHit, :1270-1284 */
    int any = 0;
    *loc = 0;
    for (uint32_t i = 0; i < s->n_lookups; i++) {
        lookup_t *l = &s->lk[i];
        if (l->n_hits > 0 && l->hits[0] - l->seed_offset > *loc) {
            s->most_recent = *loc = l->hits[0] - l->seed_offset;
            *seed_off = l->seed_offset;
            any = 1;
        }
    }
    return any;
}

static int hs_next_lower(hit_set *s, uint32_t *loc, uint32_t *seed_off)
{ /* getNextLowerHit, :1286-1322 */
    uint32_t found = 0;
    int any = 0;
    for (uint32_t i = 0; i < s->n_lookups; i++) {
        lookup_t *l = &s->lk[i];
        if (l->cur != l->n_hits && l->hits[l->cur] - l->seed_offset == s->most_recent) l->cur++;
        if (l->cur != l->n_hits) {
            if (found < l->hits[l->cur] - l->seed_offset && l->hits[l->cur] >= l->seed_offset) {
                *loc = found = l->hits[l->cur] - l->seed_offset;
                *seed_off = l->seed_offset;
                any = 1;
            }
        }
    }
    if (any) s->most_recent = found;
    return any;
}

typedef struct { /* ScoringMateCandidate, IntersectingPairedEndAligner.h:401-423 */
    double prob;
    uint32_t loc, best_possible, score, score_limit, seed_offset;
    int genome_offset;
} mate_cand;

typedef struct { /* ScoringCandidate, :425-447 */
    int next;   /* score list link */
    int anchor; /* merge anchor index or -1 */
    uint32_t mate_index, loc, set_pair, seed_offset, best_possible;
} score_cand;

typedef struct { /* MergeAnchor, :364-393 */
    double prob;
    uint32_t loc_more, loc_fewer;
    int pair_score;
} merge_anchor;

typedef struct {
    const oracle_index *x;
    snapb200_paired_params p;
    uint32_t num_seeds_cmdline;
    uint32_t pool_size;
    score_cand *cands;
    mate_cand *mates[2];
    merge_anchor *anchors;
    hit_set sets[2][2];
    double seed_prob;
    base_aligner *single;
    uint32_t n_lv, n_lookups;
    int limit_hit;
} paired_aligner;

static paired_aligner *pa_create(const oracle_index *x, const snapb200_paired_params *p)
{
    paired_aligner *a = (paired_aligner *)calloc(1, sizeof(*a));
    a->x = x; a->p = *p;
    a->num_seeds_cmdline = p->num_seeds < 30 ? p->num_seeds : 30; /* MAX_MAX_SEEDS, :47 */
    uint32_t max_seeds = a->num_seeds_cmdline ? a->num_seeds_cmdline
                                              : (uint32_t)(p->max_read_size * p->seed_coverage / x->seed_len);
    uint64_t want = (uint64_t)p->max_big_hits * max_seeds * 2;
    a->pool_size = (uint32_t)(want < p->max_candidate_pool_size ? want : p->max_candidate_pool_size); /* :128 */
    a->cands = (score_cand *)malloc(sizeof(score_cand) * (a->pool_size + 1));
    for (int i = 0; i < 2; i++) a->mates[i] = (mate_cand *)malloc(sizeof(mate_cand) * (a->pool_size / 2 + 1));
    a->anchors = (merge_anchor *)malloc(sizeof(merge_anchor) * (a->pool_size + 1));
    a->seed_prob = powi_ref(1 - 0.001, (int)x->seed_len);
    snapb200_single_params sp;
    memset(&sp, 0, sizeof(sp));
    sp.max_hits = p->max_hits; sp.max_k = p->max_k; sp.max_read_size = p->max_read_size;
    sp.num_seeds = p->num_seeds; sp.seed_coverage = p->seed_coverage; sp.extra_search_depth = p->extra_search_depth;
    a->single = ba_create(x, &sp); /* ChimericPairedEndAligner.cpp:56-58 */
    return a;
}
static void pa_destroy(paired_aligner *a)
{
    ba_destroy(a->single);
    free(a->cands); free(a->mates[0]); free(a->mates[1]); free(a->anchors); free(a);
}

/* IntersectingPairedEndAligner::align, :141-753.  Returns 0 if it returned early leaving `r` untouched. */
static int pa_intersect(paired_aligner *a, const read_views *v /*[2]*/, snapb200_paired_result *r)
{
    const oracle_index *x = a->x;
    const uint32_t seed_len = x->seed_len, max_k = a->p.max_k, extra = a->p.extra_search_depth;
    const uint32_t max_spacing = a->p.max_spacing, min_spacing = a->p.min_spacing;
    uint32_t rlen[2] = {v[0].len, v[1].len};
    uint32_t max_seeds;
    if (a->num_seeds_cmdline != 0) max_seeds = a->num_seeds_cmdline;
    else max_seeds = (uint32_t)((rlen[0] > rlen[1] ? rlen[0] : rlen[1]) * a->p.seed_coverage / seed_len);
    if (rlen[0] < 50 || rlen[1] < 50) return 0;              /* :186-188 */
    if (v[0].n_count + v[1].n_count > max_k) return 0;        /* :226-228 */

    int score_list[2 * MAXK + 8];
    for (uint32_t k = 0; k <= max_k + extra; k++) score_list[k] = -1;
    uint32_t n_cands = 0, n_mates[2] = {0, 0}, n_anchors = 0;
    uint32_t popular[2] = {0, 0}, n_look[2] = {0, 0}, total_hits[2][2] = {{0, 0}, {0, 0}};
    uint32_t best_loc[2] = {0, 0}, best_score[2] = {0, 0};
    int best_dir[2] = {0, 0};
    double p_best_pair = 0, p_all_pairs = 0;
    uint32_t best_pair_score = 65536;
    uint32_t score_limit = max_k + extra;

    /* phase 1, :259-340 */
    for (int w = 0; w < 2; w++) {
        for (int d = 0; d < 2; d++) {
            hit_set *s = &a->sets[w][d];
            s->n_lookups = 0; s->cur_set = -1; s->merge_dist = max_k; /* firstInit(maxSeeds, maxK), :114 */
        }
        uint8_t used[SNAPB200_MAX_READ_LENGTH + 8];
        memset(used, 0, sizeof(used));
        uint32_t next = 0, wrap = 0, n_possible = rlen[w] - seed_len + 1;
        int begins[2] = {1, 1};
        while (n_look[w] < n_possible && n_look[w] < max_seeds) {
            if (next >= n_possible) {
                wrap++;
                begins[0] = begins[1] = 1;
                if (wrap >= seed_len) break;
                next = wrapped_seed(seed_len, wrap);
            }
            while (next < n_possible && used[next]) next++;
            if (next >= n_possible) continue;
            used[next] = 1;
            uint64_t sf, sr;
            if (!pack_seed(v[w].data[0] + next, seed_len, &sf, &sr)) { next++; continue; } /* :296-302 */
            hit_list hl[2];
            lookup_seed(x, sf, sr, hl, NULL);
            n_look[w]++;
            for (int d = 0; d < 2; d++) {
                uint32_t offset = d == 0 ? next : rlen[w] - seed_len - next;
                if (hl[d].n < a->p.max_big_hits) {
                    total_hits[w][d] += hl[d].n;
                    if (a->sets[w][d].n_lookups >= MAX_LOOKUPS) { a->limit_hit = 1; return 0; }
                    hs_record(&a->sets[w][d], offset, hl[d].n, hl[d].hits, begins[d]);
                    begins[d] = 0;
                } else {
                    popular[w]++;
                }
            }
            if ((max_seeds - n_look[w] + 1) * seed_len + next < n_possible) /* :333-338 */
                next += (n_possible + next) / (max_seeds - n_look[w] + 1);
            else
                next += seed_len;
        }
    }
    a->n_lookups += n_look[0] + n_look[1];
    int more = (total_hits[0][0] + total_hits[0][1] > total_hits[1][0] + total_hits[1][1]) ? 0 : 1; /* :342 */
    int fewer = 1 - more;
    static const int set_dir[2][2] = {{0, 1}, {1, 0}}; /* :351 */

    /* phase 2, :359-511 */
    uint32_t max_used_list = 0;
    for (int sp = 0; sp < 2; sp++) {
        hit_set *set[2] = {&a->sets[0][set_dir[sp][0]], &a->sets[1][set_dir[sp][1]]};
        uint32_t f_loc, f_off = 0, m_loc, m_off = 0;
        int out_of_more = 0;
        if (!hs_first(set[fewer], &f_loc, &f_off)) continue;
        m_loc = INVALID_LOC;
        for (;;) {
            if (m_loc > f_loc + max_spacing) {
                if (!hs_next_le(set[more], f_loc + max_spacing, &m_loc, &m_off)) break;
            }
            if (m_loc + max_spacing < f_loc &&
                (n_mates[sp] == 0 || !is_within(a->mates[sp][n_mates[sp] - 1].loc, f_loc, max_spacing))) {
                if (!hs_next_le(set[fewer], m_loc + max_spacing, &f_loc, &f_off)) break;
                continue;
            }
            while (m_loc + max_spacing >= f_loc && !out_of_more) {
                uint32_t bp = hs_best_possible(set[more]);
                if (n_mates[sp] >= a->pool_size / 2) { a->limit_hit = 1; return 0; } /* reference soft_exits, :436-439 */
                mate_cand *m = &a->mates[sp][n_mates[sp]++];
                m->loc = m_loc; m->best_possible = bp; m->seed_offset = m_off;
                m->score = (uint32_t)-2; m->score_limit = (uint32_t)-1; m->prob = 0; m->genome_offset = 0;
                if (!hs_next_lower(set[more], &m_loc, &m_off)) {
                    m_loc = 0;
                    out_of_more = 1;
                    break;
                }
            }
            uint32_t bp_fewer = hs_best_possible(set[fewer]);
            uint32_t low_mate = max_k + extra;
            for (int i = (int)n_mates[sp] - 1; i >= 0; i--) {
                if (a->mates[sp][i].loc > f_loc + max_spacing) break;
                if (a->mates[sp][i].best_possible < low_mate) low_mate = a->mates[sp][i].best_possible;
            }
            if (low_mate + bp_fewer <= max_k + extra) {
                if (n_cands >= a->pool_size) { a->limit_hit = 1; return 0; } /* :482-485 */
                score_cand *c = &a->cands[n_cands];
                c->loc = f_loc; c->set_pair = (uint32_t)sp; c->mate_index = n_mates[sp] - 1; c->seed_offset = f_off;
                c->best_possible = bp_fewer; c->next = score_list[low_mate + bp_fewer]; c->anchor = -1;
                score_list[low_mate + bp_fewer] = (int)n_cands;
                n_cands++;
                if (low_mate + bp_fewer > max_used_list) max_used_list = low_mate + bp_fewer;
            }
            if (!hs_next_lower(set[fewer], &f_loc, &f_off)) break;
        }
    }

    /* phase 3, :516-720 */
    uint32_t list = 0;
    score_limit = max_k + extra;
    int stop = 0;
    while (!stop && list <= max_used_list && list <= score_limit) {
        if (score_list[list] < 0) { list++; continue; }
        int ci = score_list[list];
        score_cand *c = &a->cands[ci];
        double f_prob;
        int f_off;
        a->n_lv++;
        int fs = score_location(x, &v[fewer], set_dir[c->set_pair][fewer], c->loc, c->seed_offset, (int)score_limit, 0,
                                a->seed_prob, &f_prob, &f_off);
        if (fs != -1) {
            uint32_t f_score = (uint32_t)fs;
            uint32_t mi = c->mate_index;
            for (;;) {
                mate_cand *m = &a->mates[c->set_pair][mi];
                if (!is_within(m->loc, c->loc, min_spacing) && m->best_possible <= score_limit - f_score) {
                    if (m->score == (uint32_t)-2 || (m->score == (uint32_t)-1 && m->score_limit < score_limit - f_score)) {
                        a->n_lv++;
                        int ms = score_location(x, &v[more], set_dir[c->set_pair][more], m->loc, m->seed_offset,
                                                (int)(score_limit - f_score), 0, a->seed_prob, &m->prob, &m->genome_offset);
                        m->score = (uint32_t)ms;
                        m->score_limit = score_limit - f_score;
                    }
                    if (m->score != (uint32_t)-1) {
                        double pair_prob = m->prob * f_prob;
                        uint32_t pair_score = m->score + f_score;
                        uint32_t new_more = m->loc + (uint32_t)m->genome_offset, new_fewer = c->loc + (uint32_t)f_off;
                        int an = c->anchor;
                        if (an < 0) { /* :598-627 */
                            for (int j = ci - 1; j >= 0 && is_within(a->cands[j].loc, new_fewer, 50) &&
                                             a->cands[j].set_pair == c->set_pair; j--) {
                                if (a->cands[j].anchor >= 0) { c->anchor = an = a->cands[j].anchor; break; }
                            }
                            if (an < 0) {
                                /* the reference's second scan starts one above and walks DOWN (:615-619);
                                 * stepping below index 0 is out of bounds there and treated as the end here */
                                for (int j = ci + 1; j >= 0 && j < (int)n_cands && is_within(a->cands[j].loc, new_fewer, 50) &&
                                                 a->cands[j].set_pair == c->set_pair; j--) {
                                    if (a->cands[j].anchor >= 0) { c->anchor = an = a->cands[j].anchor; break; }
                                }
                            }
                        }
                        int merged;
                        double old_prob;
                        if (an < 0) {
                            if (n_anchors >= a->pool_size) { a->limit_hit = 1; return 0; }
                            merge_anchor *ma = &a->anchors[n_anchors];
                            ma->loc_more = new_more; ma->loc_fewer = new_fewer; ma->prob = pair_prob; ma->pair_score = (int)pair_score;
                            c->anchor = (int)n_anchors++;
                            merged = 0;
                            old_prob = 0;
                        } else { /* MergeAnchor::checkMerge, :1324-1371 */
                            merge_anchor *ma = &a->anchors[an];
                            uint32_t dm = ma->loc_more > new_more ? ma->loc_more - new_more : new_more - ma->loc_more;
                            uint32_t df = ma->loc_fewer > new_fewer ? ma->loc_fewer - new_fewer : new_fewer - ma->loc_fewer;
                            if (ma->loc_more == INVALID_LOC || !(dm < 50 && df < 50)) {
                                ma->loc_more = new_more; ma->loc_fewer = new_fewer; ma->prob = pair_prob; ma->pair_score = (int)pair_score;
                                old_prob = 0;
                                merged = 0;
                            } else if ((int)pair_score < ma->pair_score || ((int)pair_score == ma->pair_score && pair_prob > ma->prob)) {
                                old_prob = ma->prob;
                                ma->prob = pair_prob;
                                ma->pair_score = (int)pair_score;
                                merged = 0;
                            } else {
                                old_prob = 0;
                                merged = 1;
                            }
                        }
                        if (!merged) {
                            double t = p_all_pairs - old_prob;
                            p_all_pairs = 0 > t ? 0 : t; /* __max(0, ...) :660 */
                            if (pair_score <= max_k && (pair_score < best_pair_score ||
                                                        (pair_score == best_pair_score && pair_prob > p_best_pair))) {
                                best_pair_score = pair_score;
                                p_best_pair = pair_prob;
                                best_loc[fewer] = new_fewer; best_loc[more] = new_more;
                                best_score[fewer] = f_score; best_score[more] = m->score;
                                best_dir[fewer] = set_dir[c->set_pair][fewer]; best_dir[more] = set_dir[c->set_pair][more];
                                score_limit = best_pair_score + extra;
                            }
                            p_all_pairs += pair_prob;
                            if (p_all_pairs >= 4.9) { stop = 1; break; }
                        }
                    }
                }
                if (mi == 0 || !is_within(a->mates[c->set_pair][mi - 1].loc, c->loc, max_spacing)) break;
                mi--;
            }
        }
        if (!stop) score_list[list] = c->next;
    }

    if (best_pair_score == 65536) {
        for (int w = 0; w < 2; w++) {
            r->location[w] = INVALID_LOC; r->mapq[w] = 0; r->score[w] = -1; r->status[w] = SNAPB200_NOT_FOUND;
        }
    } else {
        for (int w = 0; w < 2; w++) {
            r->location[w] = best_loc[w];
            r->direction[w] = (uint8_t)best_dir[w];
            r->mapq[w] = compute_mapq(p_all_pairs, p_best_pair, (int)best_score[w], (int)(popular[0] + popular[1]));
            r->status[w] = r->mapq[w] > 10 ? SNAPB200_SINGLE_HIT : SNAPB200_MULTIPLE_HITS;
            r->score[w] = (int)best_score[w];
        }
    }
    r->p_all = p_all_pairs;
    r->p_best = p_best_pair;
    return 1;
}

/* ChimericPairedEndAligner::align, ChimericPairedEndAligner.cpp:74-128 */
static void pa_align(paired_aligner *a, const uint8_t *b0, const uint8_t *q0, uint32_t l0, const uint8_t *b1,
                     const uint8_t *q1, uint32_t l1, snapb200_paired_result *r)
{
    static __thread read_views v[2];
    memset(r, 0, sizeof(*r));
    r->location[0] = r->location[1] = INVALID_LOC; /* the reference leaves these uninitialised */
    r->status[0] = r->status[1] = SNAPB200_NOT_FOUND;
    if (l0 < 50 && l1 < 50) return;
    make_views(&v[0], b0, q0, l0);
    make_views(&v[1], b1, q1, l1);
    a->n_lv = 0;
    a->n_lookups = 0;
    pa_intersect(a, v, r);
    r->n_lv_calls = a->n_lv;
    r->n_lookups = a->n_lookups;
    r->from_align_together = 1;
    r->aligned_as_pair = 1;
    if (a->p.force_spacing) {
        if (r->status[0] == SNAPB200_NOT_FOUND) r->from_align_together = 0;
        return;
    }
    if (r->status[0] != SNAPB200_NOT_FOUND && r->status[1] != SNAPB200_NOT_FOUND) return;
    const uint8_t *b[2] = {b0, b1}, *q[2] = {q0, q1};
    uint32_t l[2] = {l0, l1};
    for (int e = 0; e < 2; e++) {
        snapb200_single_result sr;
        ba_align(a->single, b[e], q[e], l[e], &sr, NULL, NULL, NULL, NULL);
        r->status[e] = sr.status;
        r->location[e] = sr.location;
        r->direction[e] = sr.direction;
        r->score[e] = sr.score;
        r->mapq[e] = sr.mapq / 4;
    }
    r->from_align_together = 0;
    r->aligned_as_pair = 0;
}

int oracle_paired_batch(void *h, const snapb200_paired_params *p, const snapb200_read_batch *r0,
                        const snapb200_read_batch *r1, snapb200_paired_result *res)
{
    init_tables();
    paired_aligner *a = pa_create((const oracle_index *)h, p);
    for (uint32_t i = 0; i < r0->n; i++) {
        uint32_t o0 = r0->offsets[i], l0 = r0->offsets[i + 1] - o0;
        uint32_t o1 = r1->offsets[i], l1 = r1->offsets[i + 1] - o1;
        pa_align(a, r0->bases + o0, r0->quals + o0, l0, r1->bases + o1, r1->quals + o1, l1, &res[i]);
    }
    int lim = a->limit_hit;
    pa_destroy(a);
    return lim ? SNAPB200_ERR_LIMIT : 0;
}

/* SAMFormat::computeCigarString's aligner call, SAM.cpp:1159-1189 */
int oracle_cigar_batch(void *h, const snapb200_read_batch *reads, const uint32_t *locations, const uint8_t *directions,
                       int use_m, char *cigars, uint32_t cigar_stride, int32_t *edit_distance)
{
    const oracle_index *x = (const oracle_index *)h;
    uint8_t pat[SNAPB200_MAX_READ_LENGTH + 8];
    for (uint32_t i = 0; i < reads->n; i++) {
        uint32_t off = reads->offsets[i], len = reads->offsets[i + 1] - off;
        char *out = cigars + (size_t)i * cigar_stride;
        memset(out, 0, cigar_stride);
        if (locations[i] == INVALID_LOC || !genome_substring_ok(x, locations[i], len)) { edit_distance[i] = -3; continue; }
        if (directions[i] == SNAPB200_RC) {
            for (uint32_t q = 0; q < len; q++) pat[q] = rc_base(reads->bases[off + len - 1 - q]);
        } else {
            memcpy(pat, reads->bases + off, len);
        }
        lv_strings s;
        s.p = pat; s.plen = (int)len; s.t = x->bases + locations[i]; s.tlen = (int)len; s.dir = 1;
        s.t_lo = -(int)locations[i] - GENOME_PAD; s.t_hi = (int)((int64_t)x->n_bases + GENOME_PAD - locations[i]);
        edit_distance[i] = lv_cigar(&s, MAXK - 1, out, (int)cigar_stride, use_m);
    }
    return 0;
}

/* ProbabilityDistance::compute (SNAPLib/ProbabilityDistance.cpp:53-135) with the constructor's tables (:18-50).  The class is
 * constructed by every BaseAligner (BaseAligner.cpp:98) but `compute` is called by no aligner -- the match probability that feeds
 * MAPQ comes from LandauVishkin -- so there is no device counterpart; this restatement exists so that the reference's 16 KATs
 * (tests/ProbabilityDistanceTest.cpp:15-70) pin it and the 1e-6 agreement north_star asks for is checked against the compiled
 * reference (tests/test_oracle.py).  reference / read / quality as there; `reference` must be readable from -max_shift to
 * read_len + max_shift (the reference's own tests read outside their literals).  Three-state log-probability DP over
 * (read position, shift): NO_GAP / READ_GAP / REF_GAP. */
#define PD_MAX_SHIFT 20
#define PD_NO_PROB (-1000000.0)
static double pd_max3(double a, double b, double c) { double m = a > b ? a : b; return m > c ? m : c; }

int oracle_probability_distance(double snp_prob, double gap_open_prob, double gap_extension_prob, const char *reference, const char *read,
                                const char *quality, int read_len, int max_start_shift, int max_shift, double *match_probability)
{
    if (read_len < 0 || read_len > SNAPB200_MAX_READ_LENGTH || max_shift >= PD_MAX_SHIFT || max_start_shift > max_shift) return -1;
    const double gap_open = log(gap_open_prob), gap_ext = log(gap_extension_prob);
    double match_lp[256], mismatch_lp[256];
    for (int q = 0; q < 256; q++) {
        if (q < 33) { match_lp[q] = PD_NO_PROB; mismatch_lp[q] = PD_NO_PROB; continue; }
        const double error_prob = pow(10.0, -(q - 33) / 10.0);
        const double match_prob = (1.0 - error_prob) * (1.0 - snp_prob);
        match_lp[q] = log(match_prob);
        mismatch_lp[q] = log(1.0 - match_prob);
    }
    const int W = 2 * PD_MAX_SHIFT + 1;
    double (*d)[3] = (double (*)[3])malloc(sizeof(double) * 3 * W * (size_t)(read_len + 1));
    if (!d) return -1;
#define PD(r, s, g) d[(size_t)(r) * W + PD_MAX_SHIFT + (s)][g]
    for (int s = -max_shift - 1; s <= max_shift + 1; s++) {
        PD(0, s, 1) = PD_NO_PROB;
        PD(0, s, 2) = PD_NO_PROB;
        PD(0, s, 0) = (s < -max_start_shift || s > max_start_shift) ? PD_NO_PROB : log(1.0);
    }
    for (int r = 1; r <= read_len; r++) {
        for (int g = 0; g < 3; g++) { PD(r, -max_shift - 1, g) = PD_NO_PROB; PD(r, max_shift + 1, g) = PD_NO_PROB; }
        for (int s = -max_shift; s <= max_shift; s++) {
            const double base = (read[r - 1] == reference[r - 1 + s]) ? match_lp[(unsigned char)quality[r - 1]] : mismatch_lp[(unsigned char)quality[r - 1]];
            PD(r, s, 0) = pd_max3(PD(r - 1, s, 0) + base, PD(r - 1, s, 2) + base, PD(r - 1, s, 1) + base);
            PD(r, s, 1) = pd_max3(PD(r - 1, s + 1, 0) + gap_open, PD(r - 1, s + 1, 2) + gap_open, PD(r - 1, s + 1, 1) + gap_ext);
            PD(r, s, 2) = pd_max3(PD(r, s - 1, 0) + gap_open, PD(r, s - 1, 2) + gap_ext, PD(r, s - 1, 1) + gap_open);
        }
    }
    double best = PD_NO_PROB;
    for (int s = -max_shift; s <= max_shift; s++)
        for (int g = 0; g < 3; g++) if (PD(read_len, s, g) > best) best = PD(read_len, s, g);
#undef PD
    free(d);
    *match_probability = exp(best);
    return 5;  /* "a somewhat arbitrary score" (:133) */
}
