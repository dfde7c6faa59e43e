// lv.cuh -- warp-cooperative Landau-Vishkin bounded edit distance (score+probability, and CIGAR).
//
// Replaces LandauVishkin<+1/-1>::computeEditDistance (SNAPLib/LandauVishkin.h:211-455) and
// LandauVishkinWithCigar::computeEditDistance (SNAPLib/LandauVishkin.cpp:252-535).
//
// Mapping to the hardware: the reference fills L[e][d] one diagonal at a time in the order 0,+1,-1,...
// and returns at the first diagonal that reaches the end of the pattern.  Cells of one row depend only on
// the previous row, so here the 2e+1 diagonals of row e are extended by different lanes at once and the
// winner is the reaching diagonal with the smallest rank in the reference's visiting order -- the same
// answer, one row per step instead of one cell per step.  Row 0 (the long exact-match run) is compared 32
// bytes per step with a ballot.  L lives in shared memory as a triangular int16 table (961 cells for
// k <= 30); the action matrix A is not stored: an action is a pure function of three cells of the row
// above and is recomputed during the backtrace, which the leader lane runs (it is a dependent chain).
// Both strings are staged in shared memory by the caller, so the inner loops never touch HBM.
#pragma once
#include "common.cuh"

#define LV_CELLS (MAXK * MAXK)  // rows 0..30, row e has 2e+1 cells starting at e*e

struct LvStr {
    const uint8_t *p;  // pattern(i) = p[i*ps] for 0 <= i < plen
    int ps, plen;
    const uint8_t *t;  // text(i) = t[i*ts] for t_lo <= i < t_hi (bytes that really exist)
    int ts, tlen;      // tlen: the textLen the reference is given
    int t_lo, t_hi;
};

// Bytes outside the pattern are 0x00 and bytes outside the readable text are 0x01: they never match, which
// is what the reference's peek-then-clamp (LandauVishkin.h:325-354) amounts to whenever it is defined.
__device__ __forceinline__ int lv_pat(const LvStr &s, int i) { return ((unsigned)i < (unsigned)s.plen) ? s.p[i * s.ps] : 0x00; }
__device__ __forceinline__ int lv_txt(const LvStr &s, int i) { return (i >= s.t_lo && i < s.t_hi) ? s.t[i * s.ts] : 0x01; }

__device__ __forceinline__ int lv_get(const int16_t *L, int e, int d)
{  // cells with |d| > e are never written by the reference and keep their constructor value -2
    return (d >= -e && d <= e) ? (int)L[e * e + d + e] : -2;
}

// One cell, first part: best of (substitution, deletion, insertion) from the row above, then up to LV_QUICK positions of
// extension along the diagonal by this lane alone.  *more: the diagonal is still matching at the returned position and has
// not reached *dend -- the long runs (there are one or two per row) are finished by the whole warp, see lv_rows.
#define LV_QUICK 4
__device__ __forceinline__ int lv_cell_start(const LvStr &s, const int16_t *L, int e, int d, int *dend, bool *more)
{
    int best = lv_get(L, e - 1, d) + 1;
    int left = lv_get(L, e - 1, d - 1);
    if (left > best) best = left;
    int right = lv_get(L, e - 1, d + 1) + 1;
    if (right > best) best = right;
    *more = false;
    *dend = 0;
    if (lv_pat(s, best) == lv_txt(s, d + best)) {
        const int de = min(s.plen, s.tlen - d);
        *dend = de;
        if (best < de) {
            const int lim = min(de, best + LV_QUICK);
            #pragma unroll 1
            do { best++; } while (best < lim && lv_pat(s, best) == lv_txt(s, d + best));
            *more = best == lim && best < de && lv_pat(s, best) == lv_txt(s, d + best);
        } else {
            best = de;  // the reference's 8-byte loop clamps to `end` even when it starts beyond it
        }
    }
    return best;
}

// 'X','D','I' chosen at cell (e,d): recomputed from the row above (ties: X, then D, then I; LandauVishkin.h:312-323)
__device__ __forceinline__ char lv_action(const int16_t *L, int e, int d)
{
    int best = lv_get(L, e - 1, d) + 1;
    char a = 'X';
    int left = lv_get(L, e - 1, d - 1);
    if (left > best) { best = left; a = 'D'; }
    int right = lv_get(L, e - 1, d + 1) + 1;
    if (right > best) a = 'I';
    return a;
}

// Row 0: length of the common prefix, 32 positions per step.
__device__ __forceinline__ int lv_row0(const LvStr &s, int end)
{
    int lane = lane_id();
    #pragma unroll 1
    for (int base = 0; base < end; base += 32) {
        int i = base + lane;
        bool mism = (i < end) && (lv_pat(s, i) != lv_txt(s, i));
        unsigned b = __ballot_sync(FULL_MASK, mism);
        if (b) return base + __ffs(b) - 1;
    }
    return end;
}

// rank of diagonal d in the visiting order 0,+1,-1,+2,-2,... (score) or 0,-1,+1,-2,+2,... (CIGAR)
__device__ __forceinline__ int lv_rank_score(int d) { return d == 0 ? 0 : (d > 0 ? 2 * d - 1 : -2 * d); }
__device__ __forceinline__ int lv_unrank_score(int r) { return r == 0 ? 0 : ((r & 1) ? (r + 1) / 2 : -(r / 2)); }
__device__ __forceinline__ int lv_rank_cigar(int d) { return d == 0 ? 0 : (d < 0 ? -2 * d - 1 : 2 * d); }
__device__ __forceinline__ int lv_unrank_cigar(int r) { return r == 0 ? 0 : ((r & 1) ? -((r + 1) / 2) : r / 2); }

// Fills rows 1..k until some diagonal reaches plen.  Returns e (and the winning diagonal) or -1.
template <bool CIGAR_ORDER>
__device__ __forceinline__ int lv_rows(const LvStr &s, int16_t *L, int k, int *win_d)
{
    const int lane = lane_id();
    #pragma unroll 1
    for (int e = 1; e <= k; e++) {
        int found = 0x7fffffff;
        #pragma unroll 1
        for (int idx0 = 0; idx0 < 2 * e + 1; idx0 += 32) {
            const int idx = idx0 + lane, d = idx - e;
            const bool mine = idx < 2 * e + 1;
            int best = 0, dend = 0;
            bool more = false;
            if (mine) best = lv_cell_start(s, L, e, d, &dend, &more);
            // Diagonals that are still matching: one at a time, the whole warp compares 32 positions per step.  A lane on
            // its own would walk a 100-base exact run in 100 steps while the other lanes of the row wait for it.
            unsigned pend = __ballot_sync(FULL_MASK, more);
            #pragma unroll 1
            while (pend) {
                const int src = __ffs((int)pend) - 1;
                const int dd = __shfl_sync(FULL_MASK, d, src), de = __shfl_sync(FULL_MASK, dend, src);
                int pos = __shfl_sync(FULL_MASK, best, src);
                #pragma unroll 1
                for (;;) {
                    const int i = pos + lane;
                    const bool stop = i >= de || lv_pat(s, i) != lv_txt(s, dd + i);
                    const unsigned bm = __ballot_sync(FULL_MASK, stop);
                    if (bm) { pos += __ffs((int)bm) - 1; break; }
                    pos += 32;
                }
                if (lane == src) best = pos;
                pend &= pend - 1;
            }
            if (mine) {
                L[e * e + idx] = (int16_t)best;
                if (best == s.plen) found = min(found, CIGAR_ORDER ? lv_rank_cigar(d) : lv_rank_score(d));
            }
        }
        __syncwarp();
        int win = __reduce_min_sync(FULL_MASK, found);
        if (win != 0x7fffffff) {
            *win_d = CIGAR_ORDER ? lv_unrank_cigar(win) : lv_unrank_score(win);
            return e;
        }
    }
    return -1;
}

// LandauVishkin<DIR>::computeEditDistance.  q: quality(i) = q[i*qs] or NULL.  All lanes return the same
// values.  L: LV_CELLS int16 in shared memory private to this warp.
__device__ __noinline__ int lv_score_warp(const LvStr &s_in, const uint8_t *q, int qs, int k, int ix_slot, int16_t *L,
                             double *match_prob, int *net_indel)
{
    const DevIndex &ix = c_index[ix_slot];  // only the probability tables are used
    // a private copy: the caller's struct sits on its stack, and through the reference every field would be re-read from
    // local memory after each store to L (ncu: lv_pat/lv_txt were 45 % of the kernel's local-memory instructions)
    const LvStr s = {s_in.p, s_in.ps, s_in.plen, s_in.t, s_in.ts, s_in.tlen, s_in.t_lo, s_in.t_hi};
    const int lane = lane_id();
    *net_indel = 0;
    *match_prob = 0.0;
    if (k > MAXK - 1) k = MAXK - 1;
    const int plen = s.plen;
    int end = min(plen, s.tlen);
    int l0 = lv_row0(s, end);
    if (l0 == end) {  // LandauVishkin.h:290-305
        int result = plen > end ? plen - end : 0;
        if (q) *match_prob = ix.perfect[plen];
        return result > k ? -1 : result;
    }
    if (lane == 0) L[0] = (int16_t)l0;
    __syncwarp();
    int d = 0;
    int e = lv_rows<false>(s, L, k, &d);
    if (e < 0) return -1;
    if (!q) return e;
    // backtrace (LandauVishkin.h:379-431): a dependent chain, run by the leader lane
    double prob = 1.0;
    int indel = 0;
    if (lane == 0) {
        char act[MAXK + 1];
        short matched[MAXK + 1];
        int cur_d = d;
        #pragma unroll 1
        for (int ce = e; ce >= 1; ce--) {
            char a = lv_action(L, ce, cur_d);
            int here = (ce == e) ? plen : lv_get(L, ce, cur_d);
            act[ce] = a;
            if (a == 'I') {
                matched[ce] = (short)(here - lv_get(L, ce - 1, cur_d + 1) - 1);
                cur_d++;
            } else if (a == 'D') {
                matched[ce] = (short)(here - lv_get(L, ce - 1, cur_d - 1));
                cur_d--;
            } else {
                matched[ce] = (short)(here - lv_get(L, ce - 1, cur_d) - 1);
            }
        }
        int ce = 1;
        int offset = L[0];
        #pragma unroll 1
        while (ce <= e) {
            char a = act[ce];
            int count = 1;
            #pragma unroll 1
            while (ce + 1 <= e && matched[ce] == 0 && act[ce + 1] == a) { count++; ce++; }
            if (a == 'I') {
                prob *= ix.indel[count];
                offset += count;
                indel += count;
            } else if (a == 'D') {
                prob *= ix.indel[count];
                offset -= count;
                indel -= count;
            } else {
                #pragma unroll 1
                for (int i = 0; i < count; i++) {
                    int qi = min(plen - 1, max(offset, 0));
                    prob *= ix.phred[q[qi * qs]];
                    offset++;
                }
            }
            offset += matched[ce];
            ce++;
        }
        prob *= ix.perfect[plen - e];
    }
    *match_prob = shfl_f64(prob, 0);
    *net_indel = __shfl_sync(FULL_MASK, indel, 0);
    return e;
}

// ---- CIGAR ----------------------------------------------------------------------------------------------
struct CigarOut { char *buf; int left; };

// writeCigar, COMPACT_CIGAR_STRING case (SNAPLib/LandauVishkin.cpp:27-64): "%d%c"; false if it does not fit.
__device__ inline bool cigar_put(CigarOut &o, int count, char code)
{
    if (count <= 0) return true;
    if (o.left == 0) return false;
    char tmp[12];
    int n = 0;
    int c = count;
    #pragma unroll 1
    while (c > 0) { tmp[n++] = (char)('0' + c % 10); c /= 10; }
    int w = n + 1;
    if (w > o.left - 1) return false;
    #pragma unroll 1
    for (int i = 0; i < n; i++) o.buf[i] = tmp[n - 1 - i];
    o.buf[n] = code;
    o.buf[w] = 0;
    o.buf += w;
    o.left -= w;
    return true;
}

// LandauVishkinWithCigar::computeEditDistance, COMPACT_CIGAR_STRING.  The leader lane writes the string.
__device__ int lv_cigar_warp(const LvStr &s_in, int k, int16_t *L, char *cigar, int cigar_len, bool use_m)
{
    const LvStr s = {s_in.p, s_in.ps, s_in.plen, s_in.t, s_in.ts, s_in.tlen, s_in.t_lo, s_in.t_hi};
    const int lane = lane_id();
    const int plen = s.plen;
    int end = min(plen, s.tlen);
    int l0 = lv_row0(s, end);
    int rc = 0;
    if (l0 == end) {  // LandauVishkin.cpp:284-306
        if (lane == 0) {
            CigarOut o = {cigar, cigar_len};
            if (use_m) {
                if (!cigar_put(o, plen, 'M')) rc = -2;
            } else {
                if (!cigar_put(o, end, '=')) rc = -2;
                else if (plen > end && !cigar_put(o, plen - end, 'X')) rc = -2;
            }
        }
        return __shfl_sync(FULL_MASK, rc, 0);
    }
    if (lane == 0) L[0] = (int16_t)l0;
    __syncwarp();
    int d = 0;
    int e = lv_rows<true>(s, L, k, &d);
    if (e < 0) return -1;
    // can e plain mismatches explain it?  (LandauVishkin.cpp:357-366)
    int straight = 0;
    #pragma unroll 1
    for (int base = 0; base < end; base += 32) {
        int i = base + lane;
        bool mism = (i < end) && (lv_pat(s, i) != lv_txt(s, i));
        straight += __popc(__ballot_sync(FULL_MASK, mism));
    }
    straight += plen - end;
    // The mismatch-only case walks every base (LandauVishkin.cpp:368-395).  The comparisons are done by the whole warp, 32
    // positions per ballot, into the L table (dead once e and d are known in this case); the leader then only visits the
    // positions where match turns into mismatch or back.
    uint32_t *eq_words = (uint32_t *)L;
    if (straight == e && !use_m) {
        #pragma unroll 1
        for (int base = 0; base < end; base += 32) {
            const int i = base + lane;
            const unsigned m = __ballot_sync(FULL_MASK, i < end && lv_pat(s, i) == lv_txt(s, i));
            if (lane == 0) eq_words[base >> 5] = m;
        }
        __syncwarp();
    }
    if (lane == 0) {
        CigarOut o = {cigar, cigar_len};
        bool ok = true;
        if (straight == e) {
            if (use_m) {
                ok = cigar_put(o, plen, 'M');
            } else {
                int start = 0;
                bool matching = (eq_words[0] & 1u) != 0;
                #pragma unroll 1
                for (int base = 0; base < end && ok; base += 32) {
                    const int valid = min(32, end - base);
                    const uint32_t vmask = valid == 32 ? 0xffffffffu : ((1u << valid) - 1u);
                    const uint32_t w = eq_words[base >> 5];
                    uint32_t todo = vmask;  // positions of this word not yet passed
                    #pragma unroll 1
                    while (ok) {
                        const uint32_t x = (matching ? ~w : w) & todo;  // positions whose state differs from the current run's
                        if (!x) break;
                        const int j = __ffs((int)x) - 1;
                        const int i = base + j;
                        ok = cigar_put(o, i - start, matching ? '=' : 'X');
                        matching = !matching;
                        start = i;
                        todo = j == 31 ? 0u : (vmask & ~((2u << j) - 1u));
                    }
                }
                if (ok && plen > start) {
                    if (!matching) {
                        ok = cigar_put(o, plen - start, 'X');
                    } else {
                        ok = cigar_put(o, end - start, '=');
                        if (ok && plen > end) ok = cigar_put(o, plen - end, 'X');
                    }
                }
            }
        } else {
            // backtrace, LandauVishkin.cpp:441-531
            char act[MAXK + 1];
            short matched[MAXK + 1];
            int cur_d = d;
            #pragma unroll 1
            for (int ce = e; ce >= 1; ce--) {
                char a = lv_action(L, ce, cur_d);
                act[ce] = a;
                int here = lv_get(L, ce, cur_d);
                if (a == 'I') {
                    matched[ce] = (short)(here - lv_get(L, ce - 1, cur_d + 1) - 1);
                    cur_d++;
                } else if (a == 'D') {
                    matched[ce] = (short)(here - lv_get(L, ce - 1, cur_d - 1));
                    cur_d--;
                } else {
                    matched[ce] = (short)(here - lv_get(L, ce - 1, cur_d) - 1);
                }
            }
            int acc_m = 0;
            if (use_m) acc_m = L[0];
            else if (L[0] > 0) ok = cigar_put(o, L[0], '=');
            int ce = 1;
            #pragma unroll 1
            while (ce <= e && ok) {
                char a = act[ce];
                int count = 1;
                #pragma unroll 1
                while (ce + 1 <= e && matched[ce] == 0 && act[ce + 1] == a) { count++; ce++; }
                if (use_m) {
                    if (a == 'X') {
                        acc_m += count;
                    } else {
                        if (acc_m != 0) { ok = cigar_put(o, acc_m, 'M'); acc_m = 0; }
                        if (ok) ok = cigar_put(o, count, a);
                    }
                } else {
                    ok = cigar_put(o, count, a);
                }
                if (ok && matched[ce] > 0) {
                    if (use_m) acc_m += matched[ce];
                    else ok = cigar_put(o, matched[ce], '=');
                }
                ce++;
            }
            if (ok && use_m && acc_m != 0) ok = cigar_put(o, acc_m, 'M');
        }
        rc = ok ? e : -2;
    }
    return __shfl_sync(FULL_MASK, rc, 0);
}

// =================================================================================================================
// Lane-level Landau-Vishkin: 32 candidates per warp, one per lane.
//
// The warp-cooperative routine above spends a whole warp on one candidate, which is the right trade for the common
// pair (two or three candidates) but not for reads from repeat families, where one pair scores thousands of
// candidates and the kernel becomes issue bound.  For those, every lane runs the *same* algorithm on its own
// candidate: cells are visited in the reference's order (so the first diagonal to reach the end of the pattern is
// the reference's), strings are compared four bytes per step (XOR + find-first-set, as the reference does with
// eight), the pattern comes from the read staged in shared memory and the text straight from the genome in HBM
// through L1 (a candidate's window is two or three 128-byte lines, touched by one lane only).  The per-lane L table
// is interleaved in shared memory (cell c of lane l at L[c*32+l]) so that lanes walking the table in lockstep never
// conflict.  Preconditions (checked by the caller): k <= kl, textLen >= patternLen + k (always true for
// windows inside the genome), and the 4-byte over-reads stay inside readable memory.
// =================================================================================================================
// The largest score limit handled in lane mode, `kl`, is a property of the launch: max_k + extra_search_depth of the
// run (17 with the paired defaults, 22 for the -d 20 stress configuration), so the shared-memory rows are as wide as the
// run can need and no wider.
// A cell of the rolling rows holds -2 .. pattern length.  One byte (biased by 2) is enough for the strings lane mode is used on
// (reads of at most LANE_MAX_READ bases; longer reads are scored in warp mode), and halves the shared memory of a warp's rows --
// which is what decides how many warps an SM holds.  -DLANE_ROWS_I16 builds the 16-bit rows of the first version.
#ifdef LANE_ROWS_I16
typedef int16_t lane_cell_t;
#define LC_BIAS 0
#define LANE_MAX_READ 500
#else
typedef uint8_t lane_cell_t;
#define LC_BIAS 2
#define LANE_MAX_READ 250
#endif
#define LC_ST(v) ((lane_cell_t)((v) + LC_BIAS))
#define LC_LD(x) ((int)(x) - LC_BIAS)
__host__ __device__ inline int lane_rowp(int kl) { return 2 * kl + 3; }  // a row: diagonals -kl..kl plus one cell either side, so that a row's out-of-band neighbours exist
__host__ __device__ inline int lane_roll_cells(int kl) { return 2 * lane_rowp(kl); }          // per lane: previous row + current row (shared memory)
__host__ __device__ inline int lane_table_cells(int kl) { return (kl + 1) * (kl + 1); }       // per lane: the full triangular table (HBM scratch)

// length of the common run of pattern[pi..] and text[ti..], at most plen - pi.  Both strings are walked one aligned
// 32-bit word per step (the byte alignment of each string is constant along the run, so the funnel-shift amounts are
// too); the text word of the NEXT step is requested before the current one is compared, which takes the L1/L2 latency
// of the genome off the dependent chain.  Words up to seven bytes beyond either string are read (never compared).
template <int DIR>
__device__ __forceinline__ int lane_run(const uint8_t *p, const uint8_t *t, int pi, int ti, int plen)
{
    const int rem = plen - pi;
    if (rem <= 0) return 0;
    // first byte compared in memory order: string(i) lives at base + DIR * i; a step covers 4 bytes at [a, a+3]
    const uintptr_t pa = (uintptr_t)(DIR > 0 ? p + pi : p - pi - 3), ta = (uintptr_t)(DIR > 0 ? t + ti : t - ti - 3);
    const uint32_t *pw = (const uint32_t *)(pa & ~(uintptr_t)3), *tw = (const uint32_t *)(ta & ~(uintptr_t)3);
    const unsigned psh = (unsigned)(pa & 3) * 8, tsh = (unsigned)(ta & 3) * 8;
    uint32_t p0 = pw[0], p1 = pw[1], t0 = tw[0], t1 = tw[1];
    int n = 0;
    #pragma unroll 1
    for (;;) {
        // the words the next step needs: one further along the walk (up for DIR > 0, down for DIR < 0)
        pw += DIR; tw += DIR;
        const uint32_t pn = DIR > 0 ? pw[1] : pw[0], tn = DIR > 0 ? tw[1] : tw[0];
        uint32_t x = __funnelshift_r(p0, p1, psh) ^ __funnelshift_r(t0, t1, tsh);
        if (x) {
            if (DIR < 0) x = __byte_perm(x, 0, 0x0123);  // string order is descending addresses
            n += (__ffs((int)x) - 1) >> 3;
            break;
        }
        n += 4;
        if (n >= rem) break;
        if (DIR > 0) { p0 = p1; p1 = pn; t0 = t1; t1 = tn; } else { p1 = p0; p0 = pn; t1 = t0; t0 = tn; }
    }
    return n < rem ? n : rem;
}

// the same with the direction as a run-time value: one copy of the code for both passes (the kernel is instruction-fetch bound)
__device__ __forceinline__ int lane_run_rt(const int DIR, const uint8_t *p, const uint8_t *t, int pi, int ti, int plen)
{
    const int rem = plen - pi;
    if (rem <= 0) return 0;
    // first byte compared in memory order: string(i) lives at base + DIR * i; a step covers 4 bytes at [a, a+3]
    const uintptr_t pa = (uintptr_t)(DIR > 0 ? p + pi : p - pi - 3), ta = (uintptr_t)(DIR > 0 ? t + ti : t - ti - 3);
    const uint32_t *pw = (const uint32_t *)(pa & ~(uintptr_t)3), *tw = (const uint32_t *)(ta & ~(uintptr_t)3);
    const unsigned psh = (unsigned)(pa & 3) * 8, tsh = (unsigned)(ta & 3) * 8;
    uint32_t p0 = pw[0], p1 = pw[1], t0 = tw[0], t1 = tw[1];
    int n = 0;
    #pragma unroll 1
    for (;;) {
        // the words the next step needs: one further along the walk (up for DIR > 0, down for DIR < 0)
        pw += DIR; tw += DIR;
        // (a second word of look-ahead for the text was measured: 62.6 against 61.8 ms per million pairs, profiles/r2_ab_8.log)
        const uint32_t pn = DIR > 0 ? pw[1] : pw[0], tn = DIR > 0 ? tw[1] : tw[0];
        uint32_t x = __funnelshift_r(p0, p1, psh) ^ __funnelshift_r(t0, t1, tsh);
        if (x) {
            if (DIR < 0) x = __byte_perm(x, 0, 0x0123);  // string order is descending addresses
            n += (__ffs((int)x) - 1) >> 3;
            break;
        }
        n += 4;
        if (n >= rem) break;
        if (DIR > 0) { p0 = p1; p1 = pn; t0 = t1; t1 = tn; } else { p1 = p0; p0 = pn; t1 = t0; t0 = tn; }
    }
    return n < rem ? n : rem;
}

// full table (HBM scratch): cell (e,d) of this lane at T[(e*e+d+e)*32]; cells outside the band read as -2
__device__ __forceinline__ int lane_get(const lane_cell_t *T, int e, int d)
{
    return (d >= -e && d <= e) ? LC_LD(T[(e * e + d + e) * 32]) : -2;
}
// rolling rows (shared): row parity (e&1), diagonal d of this lane at R[((e&1)*rowp + d + kl + 1)*32]

// LandauVishkin<DIR>::computeEditDistance for this lane's candidate.  All 32 lanes must call it together (it uses
// a warp vote to stop early); `live_in` is false for lanes without a candidate.  p/t point at string index 0 and are
// walked with stride DIR; q likewise.  R = this lane's column of the rolling rows in shared memory, T = this lane's
// column of the full table in HBM scratch (written on the way, read only by the backtrace of successful lanes).
// k may differ between lanes.  Returns the score or -1.
template <int DIR>
__device__ __noinline__ int lv_lane(const uint8_t *p, int plen, const uint8_t *t, const uint8_t *q, int k, int kl, lane_cell_t *R, lane_cell_t *T, int ix_slot,
                       bool live_in, double *match_prob, int *net_indel)
{
    const int rowp = lane_rowp(kl);
    const DevIndex &ix = c_index[ix_slot];  // only the probability tables are used
    int result = -1, win_d = 0;
    *match_prob = 0.0;
    *net_indel = 0;
    bool live = live_in;
    int l0 = 0;
    if (live) {
        l0 = lane_run<DIR>(p, t, 0, 0, plen);
        R[(kl + 1) * 32] = LC_ST(l0);
        T[0] = LC_ST(l0);
        if (l0 == plen) {  // LandauVishkin.h:290-305 (text is never shorter than the pattern here)
            result = 0;
            *match_prob = ix.perfect[plen];
            live = false;
        }
    }
    const int kmax = min(k > 0 ? k : 0, kl);
    #pragma unroll 1
    for (int e = 1; e <= kl; e++) {
        if (live && e > kmax) live = false;
        if (!__any_sync(FULL_MASK, live)) break;
        // this lane's column of the previous and the current row, centred on diagonal 0
        lane_cell_t *prev = R + ((((e - 1) & 1) * rowp) + kl + 1) * 32, *cur = R + (((e & 1) * rowp) + kl + 1) * 32;
        lane_cell_t *Te = T + (e * e + e) * 32;
        if (live) {  // cells just outside the band of row e-1 read as -2 (never written by the reference); no range tests below
            prev[e * 32] = LC_ST(-2); prev[-e * 32] = LC_ST(-2); prev[(e + 1) * 32] = LC_ST(-2); prev[-(e + 1) * 32] = LC_ST(-2);
        }
        int d = 0;  // visiting order 0,+1,-1,+2,-2,...
        #pragma unroll 1
        for (int r = 0; r <= 2 * e; r++) {
            if (live) {
                int best = LC_LD(prev[d * 32]) + 1;
                const int left = LC_LD(prev[(d - 1) * 32]);
                if (left > best) best = left;
                const int right = LC_LD(prev[(d + 1) * 32]) + 1;
                if (right > best) best = right;
                if (best < plen) best += lane_run<DIR>(p, t, best, d + best, plen);
                cur[d * 32] = LC_ST(best);
                Te[d * 32] = LC_ST(best);
                if (best == plen) { result = e; win_d = d; live = false; }
            }
            d = d > 0 ? -d : 1 - d;
        }
    }
    if (result >= 1) {  // backtrace, LandauVishkin.h:379-431
        char act[MAXK + 1];
        short matched[MAXK + 1];
        int cur_d = win_d;
        #pragma unroll 1
        for (int ce = result; ce >= 1; ce--) {
            int up = lane_get(T, ce - 1, cur_d) + 1, left = lane_get(T, ce - 1, cur_d - 1), right = lane_get(T, ce - 1, cur_d + 1) + 1;
            int best = up;
            char a = 'X';
            if (left > best) { best = left; a = 'D'; }
            if (right > best) { best = right; a = 'I'; }
            const int here = (ce == result) ? plen : lane_get(T, ce, cur_d);
            act[ce] = a;
            matched[ce] = (short)(here - best);  // bases matched after the edit = L[ce][d] - (value before extension)
            cur_d += a == 'I' ? 1 : (a == 'D' ? -1 : 0);
        }
        double prob = 1.0;
        int indel = 0, ce = 1, offset = l0;
        #pragma unroll 1
        while (ce <= result) {
            const char a = act[ce];
            int count = 1;
            #pragma unroll 1
            while (ce + 1 <= result && matched[ce] == 0 && act[ce + 1] == a) { count++; ce++; }
            if (a == 'I') {
                prob *= ix.indel[count];
                offset += count;
                indel += count;
            } else if (a == 'D') {
                prob *= ix.indel[count];
                offset -= count;
                indel -= count;
            } else {
                #pragma unroll 1
                for (int i = 0; i < count; i++) {
                    int qi = min(plen - 1, max(offset, 0));
                    prob *= ix.phred[q[qi * DIR]];
                    offset++;
                }
            }
            offset += matched[ce];
            ce++;
        }
        prob *= ix.perfect[plen - result];
        *match_prob = prob;
        *net_indel = indel;
    }
    return result;
}

// one body for both directions (see lane_run_rt)
__device__ __noinline__ int lv_lane_rt(const int DIR, const uint8_t *p, int plen, const uint8_t *t, const uint8_t *q, int k, int kl, lane_cell_t *R, lane_cell_t *T, int ix_slot,
                       bool live_in, double *match_prob, int *net_indel)
{
    const int rowp = lane_rowp(kl);
    const DevIndex &ix = c_index[ix_slot];  // only the probability tables are used
    int result = -1, win_d = 0;
    *match_prob = 0.0;
    *net_indel = 0;
    bool live = live_in;
    int l0 = 0;
    if (live) {
        l0 = lane_run_rt(DIR, p, t, 0, 0, plen);
        R[(kl + 1) * 32] = LC_ST(l0);
        T[0] = LC_ST(l0);
        if (l0 == plen) {  // LandauVishkin.h:290-305 (text is never shorter than the pattern here)
            result = 0;
            *match_prob = ix.perfect[plen];
            live = false;
        }
    }
    const int kmax = min(k > 0 ? k : 0, kl);
    #pragma unroll 1
    for (int e = 1; e <= kl; e++) {
        if (live && e > kmax) live = false;
        if (!__any_sync(FULL_MASK, live)) break;
        // this lane's column of the previous and the current row, centred on diagonal 0
        lane_cell_t *prev = R + ((((e - 1) & 1) * rowp) + kl + 1) * 32, *cur = R + (((e & 1) * rowp) + kl + 1) * 32;
        lane_cell_t *Te = T + (e * e + e) * 32;
        if (live) {  // cells just outside the band of row e-1 read as -2 (never written by the reference); no range tests below
            prev[e * 32] = LC_ST(-2); prev[-e * 32] = LC_ST(-2); prev[(e + 1) * 32] = LC_ST(-2); prev[-(e + 1) * 32] = LC_ST(-2);
        }
        int d = 0;  // visiting order 0,+1,-1,+2,-2,...
        #pragma unroll 1
        for (int r = 0; r <= 2 * e; r++) {
            if (live) {
                int best = LC_LD(prev[d * 32]) + 1;
                const int left = LC_LD(prev[(d - 1) * 32]);
                if (left > best) best = left;
                const int right = LC_LD(prev[(d + 1) * 32]) + 1;
                if (right > best) best = right;
                if (best < plen) best += lane_run_rt(DIR, p, t, best, d + best, plen);
                cur[d * 32] = LC_ST(best);
                Te[d * 32] = LC_ST(best);
                if (best == plen) { result = e; win_d = d; live = false; }
            }
            d = d > 0 ? -d : 1 - d;
        }
    }
    if (result >= 1) {  // backtrace, LandauVishkin.h:379-431
        char act[MAXK + 1];
        short matched[MAXK + 1];
        int cur_d = win_d;
        #pragma unroll 1
        for (int ce = result; ce >= 1; ce--) {
            int up = lane_get(T, ce - 1, cur_d) + 1, left = lane_get(T, ce - 1, cur_d - 1), right = lane_get(T, ce - 1, cur_d + 1) + 1;
            int best = up;
            char a = 'X';
            if (left > best) { best = left; a = 'D'; }
            if (right > best) { best = right; a = 'I'; }
            const int here = (ce == result) ? plen : lane_get(T, ce, cur_d);
            act[ce] = a;
            matched[ce] = (short)(here - best);  // bases matched after the edit = L[ce][d] - (value before extension)
            cur_d += a == 'I' ? 1 : (a == 'D' ? -1 : 0);
        }
        double prob = 1.0;
        int indel = 0, ce = 1, offset = l0;
        #pragma unroll 1
        while (ce <= result) {
            const char a = act[ce];
            int count = 1;
            #pragma unroll 1
            while (ce + 1 <= result && matched[ce] == 0 && act[ce + 1] == a) { count++; ce++; }
            if (a == 'I') {
                prob *= ix.indel[count];
                offset += count;
                indel += count;
            } else if (a == 'D') {
                prob *= ix.indel[count];
                offset -= count;
                indel -= count;
            } else {
                #pragma unroll 1
                for (int i = 0; i < count; i++) {
                    int qi = min(plen - 1, max(offset, 0));
                    prob *= ix.phred[q[qi * DIR]];
                    offset++;
                }
            }
            offset += matched[ce];
            ce++;
        }
        prob *= ix.perfect[plen - result];
        *match_prob = prob;
        *net_indel = indel;
    }
    return result;
}
