// paired.cuh -- paired-end set-intersection aligner, one warp per pair.
//
// Replaces IntersectingPairedEndAligner::align / scoreLocation / HashTableHitSet / MergeAnchor
// (SNAPLib/IntersectingPairedEndAligner.cpp:141-1371).  The single-end fallback of
// ChimericPairedEndAligner::align (SNAPLib/ChimericPairedEndAligner.cpp:74-128) is a second launch of the
// single-end kernel over the pairs this kernel could not place (see snapb200.cu).
//
// Hardware mapping: phase 1 probes all seeds of both mates at once (one lane per seed); in phase 2 every lane
// owns one lookup of a hit set, so the per-step "binary search in every lookup, take the maximum" of the
// reference becomes one search per lane plus a warp max-reduction; phase 3 scores candidates with the
// warp-cooperative Landau-Vishkin of lv.cuh, the leader lane keeping the merge/MAPQ bookkeeping in the
// reference's order.
#pragma once
#include <cstddef>
#include "single.cuh"

#define SC_NONE (-3)       // "score not computed yet" in the look-ahead caches below
// Tunables (measured on C3, ms per million pairs: look-ahead rounds 1/2/3 = 117/95/96; minimum lane batch 2/3/6 = 102/95/95)
#ifndef MATE_LOOKAHEAD_ROUNDS
#define MATE_LOOKAHEAD_ROUNDS 2  // mates scored ahead per candidate of a lane-mode batch
#endif
#ifndef LANE_MIN_BATCH
#define LANE_MIN_BATCH 1  // fewer pending locations than this: the warp-cooperative LV is used.  Round 2: 1 -- every location that may be
                          // scored in lane mode is (63.5 -> 62.0 ms per million C3 pairs): the warp-mode routines stay in the binary for windows at
                          // the genome's edge, limits above lane_k and long reads, but are no longer fetched by ordinary pairs
#endif

struct __align__(8) Mate {  // ScoringMateCandidate (IntersectingPairedEndAligner.h:401-423)
    double prob;
    uint32_t loc, best_possible, score, score_limit;
    uint32_t seed_offset;
    int32_t genome_offset;
    // look-ahead cache: the true LV outcome for limit s_k (s_score = -1: distance > s_k)
    double s_prob;
    int16_t s_score;
    uint8_t s_k;
    int8_t s_off;
    uint32_t pad;
};
struct __align__(8) Cand {  // ScoringCandidate (:425-447); `list` = the score list it is on (bestPossibleScore of the pair)
    int32_t list, anchor;
    uint32_t mate_index, loc;
    uint16_t seed_offset;
    uint8_t set_pair, best_possible;
    // look-ahead cache of the fewer end's score (see phase 3)
    int16_t c_score;
    uint8_t c_k;
    int8_t c_off;
    double c_prob;
};
// is the outcome of scoring this mate with limit `limit` already determined by the look-ahead cache?
__device__ __forceinline__ bool mate_known(const Mate *m, uint32_t limit)
{
    return m->s_score >= 0 || (m->s_score == -1 && m->s_k >= limit);  // a stored distance is exact for every limit
}

struct __align__(8) Anchor {  // MergeAnchor (:364-393)
    double prob;
    uint32_t loc_more, loc_fewer;
    int32_t pair_score;
    int32_t pad;
};

struct PairedCfg {
    uint32_t max_k, num_seeds, extra, min_spacing, max_spacing, max_big_hits;
    double seed_coverage;
    uint32_t cand_cap, mate_cap, anchor_cap;  // this scratch tier
    uint32_t hard_limit;                       // 1: caps are the reference's pool sizes (overflow = its soft_exit)
    uint32_t lane_k;                           // largest score limit scored in lane mode = max_k + extra (sizes the LV rows)
    uint32_t lane_gate;                        // a limit is scored in lane mode iff limit < lane_gate: lane_k + 1, or 0 when the reads are too long for the one-byte rows
    uint32_t rl;
};

struct PairedScratch {
    Cand *cands;
    Mate *mates;          // two arrays of mate_cap entries, one per set pair
    uint32_t mate_cap;
    __device__ __forceinline__ Mate *mates_of(uint32_t sp) const { return mates + (size_t)sp * mate_cap; }
    Anchor *anchors;
    lane_cell_t *lane_table;  // lane_table_cells(cfg.lane_k) * 32 cells: the full L tables of a lane-mode batch
    uint32_t *order;      // cand_cap entries: candidate indices in phase 3's visiting order
};

#define STATUS_LIMIT 0xfd  // the reference's candidate pools would have overflowed (it exits)

// phase 1 staging (raw lookup results per scheduled seed); lives in the Landau-Vishkin buffer, which is idle until phase 3
struct Phase1Sm {
    unsigned long long raw_hits[2][MAX_LOOKUPS];
    uint32_t raw_n[2][MAX_LOOKUPS];
    uint32_t raw_last[2][MAX_LOOKUPS];  // last word of each hit list (the trim test of recordLookup), loaded by the probing lane
    uint16_t sched_off[MAX_LOOKUPS];
    uint8_t sched_wrap[MAX_LOOKUPS];
    uint32_t used[16];
};

struct PairedSm {
    // hit sets [read][dir], filled by the leader in phase 1 (HashTableHitSet::recordLookup)
    // hit list of lookup k: overflow-table word offset of its first hit, or -- for a seed with exactly one hit, which the
    // index stores inside the hash-table entry -- the hit itself (bit k of inline_mask set)
    uint32_t hitref[2][2][MAX_LOOKUPS];
    uint32_t inline_mask[2][2];
    uint32_t nhits[2][2][MAX_LOOKUPS];
    uint16_t seedoff[2][2][MAX_LOOKUPS];
    uint8_t setid[2][2][MAX_LOOKUPS];
    uint8_t exhausted[2][2][MAX_LOOKUPS];
    int8_t cur_set[2][2];
    uint8_t n_lookups[2][2];
    uint32_t n_sched_w[2];
    uint32_t total_hits[2][2], popular[2], n_look[2];
    uint32_t list_pos[32];   // counting sort of the candidates by score list (phase 2 -> 3)
    // The seed schedule of a read without non-ACGT bases depends only on its length (and the seed count), and the reads of a
    // batch mostly share one length: the last schedule computed is kept and reused (sc_key = length | seeds << 16).
    uint32_t sc_key[2], sc_n[2], sched_cached[2];  // one cache per mate slot
    uint16_t sc_off[2][MAX_LOOKUPS / 2];
    uint8_t sc_wrap[2][MAX_LOOKUPS / 2];
    unsigned long long acc[5];  // this warp's share of the run counters (lookups, locations scored, popular seeds, table probes, hit words)
    uint32_t ring_loc[32];   // phase 2: location and bestPossibleScore of the newest 32 mates (entry i at i & 31)
    uint32_t ring_bp[32];
    // phase 3 exchange
    double p_all, p_best, f_prob, m_prob;
    uint32_t best_pair_score, score_limit, n_cands, n_anchors, n_mates[2], pos;
    uint32_t best_loc[2], best_score[2];
    int best_dir[2];
    int act, act2, ci, stop, overflow, state, f_off, m_off, fs, ms;
    uint32_t n_batch;
    uint32_t batch_ids[32];
    uint32_t c_loc, c_seedoff, c_sp, mi, m_loc, m_seedoff, m_limit, low_mate;
    uint32_t n_lv, n_probes, n_hit_words;
    PROF(long long t_phase[12];)  // cycle accounting (-DSNAPB200_PROFILE): stage, phase1, phase2, lv, leader3, other
};

// ---- warp-parallel HashTableHitSet: lane i owns lookup i ------------------------------------------------
struct LaneLookup {
    const uint32_t *hits;  // in the overflow table; unused when the only hit is held inline
    uint32_t single;       // the inline hit
    bool inl;
    uint32_t nh, cur, so, sid;
    uint32_t cur_val, prev_val;  // hits[cur] (if cur < nh) and hits[cur-1] (if cur > 0), kept in registers
    uint32_t next_val;           // hits[cur+1] (if cur+1 < nh): requested when cur is set, so stepping down never waits for L2
    uint32_t words;  // hit-list words this lane has read (accounting for the roofline figure)
    bool act;
};

// hit i of a lane's list (an inline list has at most one hit, so i is 0 there)
__device__ __forceinline__ uint32_t hit_at(const LaneLookup &l, uint32_t i) { return l.inl ? l.single : __ldg(&l.hits[i]); }

__device__ __forceinline__ LaneLookup load_lookup(const PairedSm *sm, const uint32_t *overflow, int w, int d)
{
    LaneLookup l;
    const int lane = lane_id();
    l.act = lane < (int)sm->n_lookups[w][d];
    l.single = l.act ? sm->hitref[w][d][lane] : 0;
    l.inl = l.act && (sm->inline_mask[w][d] >> lane & 1);
    l.hits = overflow + (l.inl ? 0u : l.single);
    l.nh = l.act ? sm->nhits[w][d][lane] : 0;
    l.so = l.act ? sm->seedoff[w][d][lane] : 0;
    l.sid = l.act ? sm->setid[w][d][lane] : 0;
    l.cur = 0;
    l.words = 0;
    l.cur_val = l.nh > 0 ? hit_at(l, 0) : 0;
    l.next_val = l.nh > 1 ? hit_at(l, 1) : 0;
    l.prev_val = 0;
    l.words += l.nh > 0;
    return l;
}

__device__ __forceinline__ bool is_within(uint32_t a, uint32_t b, uint32_t dist)
{  // Util.h:538-541 with its unsigned wrap-around
    // (a <= b && a + dist >= b) || (a >= b && a <= b + dist), as one select instead of a chain of branches
    return a <= b ? (uint32_t)(a + dist) >= b : a <= (uint32_t)(b + dist);
}

// max over lanes of (ok ? val : 0) with the reference's "first strictly greater wins" tie rule; returns whether
// any lane had ok && val > 0, the winning value and that lane's seed offset.
__device__ __forceinline__ bool pick_max(bool ok, uint32_t val, uint32_t so, uint32_t *best, uint32_t *best_so)
{
    uint32_t v = ok ? val : 0;
    uint32_t m = __reduce_max_sync(FULL_MASK, v);
    if (m == 0) return false;
    unsigned who = __ballot_sync(FULL_MASK, v == m);  // m > 0, so v == m implies ok
    int src = __ffs(who) - 1;
    *best = m;
    *best_so = __shfl_sync(FULL_MASK, so, src);
    return true;
}

// getFirstHit (:1270-1284)
__device__ __forceinline__ bool hs_first(LaneLookup &l, uint32_t *most_recent, uint32_t *loc, uint32_t *so)
{
    bool ok = l.act && l.nh > 0;
    uint32_t val = ok ? l.cur_val - l.so : 0;
    *loc = 0;
    if (!pick_max(ok, val, l.so, loc, so)) return false;
    *most_recent = *loc;
    return true;
}

// getNextHitLessThanOrEqualTo, "traditional" branch (:1219-1263)
__device__ __forceinline__ bool hs_next_le(LaneLookup &l, uint32_t *most_recent, uint32_t max_loc, uint32_t *loc, uint32_t *so)
{
    bool found = false;
    uint32_t val = 0;
    if (l.act) {
        int lo = (int)l.cur, hi = (int)l.nh - 1;
        uint32_t want = max_loc + l.so;
        #pragma unroll 1
        while (lo <= hi) {
            int probe = (lo + hi) / 2;
            uint32_t h = hit_at(l, (uint32_t)probe);
            uint32_t hp = probe == 0 ? 0 : hit_at(l, (uint32_t)probe - 1);
            l.words += 2;
            if (h <= want && (probe == 0 || hp > want)) {
                found = true;
                val = h - l.so;
                l.cur = (uint32_t)probe;
                l.cur_val = h;
                l.prev_val = hp;
                l.next_val = (uint32_t)probe + 1 < l.nh ? hit_at(l, (uint32_t)probe + 1) : 0;
                break;
            }
            if (h > want) lo = probe + 1; else hi = probe - 1;
        }
        if (lo > hi) {
            l.cur = l.nh;
            l.prev_val = l.nh > 0 ? hit_at(l, l.nh - 1) : 0;
        }
    }
    if (!pick_max(found, val, l.so, loc, so)) return false;
    *most_recent = *loc;
    return true;
}

// getNextLowerHit (:1286-1322).  A lane without a lookup has nh == cur == 0 and falls through everything.
__device__ __forceinline__ bool hs_next_lower(LaneLookup &l, uint32_t *most_recent, uint32_t *loc, uint32_t *so)
{
    if (l.cur != l.nh && l.cur_val - l.so == *most_recent) {
        l.cur++;
        l.prev_val = l.cur_val;
        l.cur_val = l.next_val;
        l.words += l.cur != l.nh;
        if (l.cur + 1 < l.nh) l.next_val = __ldg(&l.hits[l.cur + 1]);  // never an inline list: those have at most one hit
    }
    const uint32_t val = l.cur_val - l.so;
    const bool ok = l.cur != l.nh && l.cur_val >= l.so;
    if (!pick_max(ok, val, l.so, loc, so)) return false;
    *most_recent = *loc;
    return true;
}

// computeBestPossibleScoreForCurrentHit (:901-929); merge_dist = maxK (firstInit(maxSeeds, maxK), :114).
// exh_mine = exhausted[l.sid] and max_exh = max over sets of exhausted[] are constant during phase 2.
__device__ __forceinline__ uint32_t hs_best_possible(const LaneLookup &l, uint32_t exh_mine, uint32_t max_exh, uint32_t most_recent,
                                                     uint32_t merge_dist)
{
    // isWithin(hit, target, mergeDist) for the current hit and the one before it.  The walk is strictly descending and
    // most_recent is the maximum over the lanes' current hits, so cur_val <= target < prev_val always: each test is one
    // subtraction and one compare (no wrap-around: locations + MAX_K stay below 2^32, checked when the index is made).
    const uint32_t target = most_recent + l.so;
    const bool close = ((l.cur != l.nh) & (target - l.cur_val <= merge_dist)) | ((l.cur != 0) & (l.prev_val - target <= merge_dist));
    const bool miss = l.act & !close;
    // misses per disjoint hit set: lanes of the same set that missed find each other with one match
    const unsigned peers = __match_any_sync(FULL_MASK, miss ? l.sid : 0xffffffffu);
    const uint32_t mine = miss ? exh_mine + (uint32_t)__popc(peers) : 0u;
    return max(max_exh, __reduce_max_sync(FULL_MASK, mine));
}

// leader: the seed schedule of one mate (IntersectingPairedEndAligner.cpp:259-339); every non-N seed is a lookup.
// Offsets go to sched_off[slot0..]; all_acgt: the read has no base that could invalidate a seed (skips the per-seed test).
__device__ __noinline__ void schedule_seeds_paired(PairedSm *sm, Phase1Sm *p1, int w, uint32_t slot0, const uint8_t *read, uint32_t len, uint32_t seed_len,
                                                      uint32_t max_seeds, bool all_acgt)
{
    const uint32_t n_possible = len - seed_len + 1;
    #pragma unroll 1
    for (int i = 0; i < 16; i++) p1->used[i] = 0;
    uint32_t next = 0, wrap = 0, n = 0;
    #pragma unroll 1
    while (n < n_possible && n < max_seeds) {
        if (next >= n_possible) {
            wrap++;
            if (wrap >= seed_len) break;
            next = wrapped_seed(seed_len, wrap);
        }
        #pragma unroll 1
        while (next < n_possible && (p1->used[next >> 5] >> (next & 31) & 1)) next++;
        if (next >= n_possible) continue;
        p1->used[next >> 5] |= 1u << (next & 31);
        if (!all_acgt) {
            bool ok = true;
            #pragma unroll 1
            for (uint32_t i = 0; i < seed_len; i++) ok &= base2(read[next + i]) >= 0;
            if (!ok) { next++; continue; }  // :296-302
        }
        p1->sched_off[slot0 + n] = (uint16_t)next;
        p1->sched_wrap[slot0 + n] = (uint8_t)wrap;
        n++;
        if ((max_seeds - n + 1) * seed_len + next < n_possible)  // :333-338 (n == countOfHashTableLookups here)
            next += (n_possible + next) / (max_seeds - n + 1);
        else
            next += seed_len;
    }
    sm->n_sched_w[w] = n;
}

// Lane mode for the candidates of phase 3: lane i scores candidate batch_ids[i] with the current limit, then the mates those
// candidates will ask about are scored ahead (see phase 3).  Out of line: one copy, called from the one place of the main body that
// needs it (with LANE_MIN_BATCH 1 every pair comes here, an ordinary pair with a batch of one candidate and then one mate).
// (arguments by value: a reference to the kernel's configuration or scratch descriptor would force them onto the stack)
__device__ __noinline__ void lane_batch_candidates(int ix_slot, int lane_k, uint32_t min_spacing, uint32_t max_spacing, Cand *cands, Mate *mates,
                                                   uint32_t mate_cap, lane_cell_t *lane_table, PairedSm *sm, const ReadView vf, const ReadView vm,
                                                   int fewer, uint32_t n_batch, int16_t *L)
{
    const int lane = lane_id();
    const int more = 1 - fewer;
    PROF(long long t_x = clock64();)
    // lane mode: lane i scores candidate batch_ids[i] with K = current limit
    PROF(long long t_y = clock64();)
    const int K = (int)sm->score_limit;
    bool act_l = (uint32_t)lane < n_batch;
    Cand *cl = act_l ? &cands[sm->batch_ids[lane]] : nullptr;
    int s = SC_NONE, off = 0;
    double pr = 0;
    const int dl = act_l ? (fewer == 0 ? (int)cl->set_pair : 1 - (int)cl->set_pair) : 0;
    score_location_lane(ix_slot, vf, dl, act_l ? cl->loc : 0, act_l ? cl->seed_offset : 0, K, lane_k, (lane_cell_t *)L + lane, lane_table + lane, act_l, &s, &pr, &off);
    if (act_l && s != SC_NONE) { cl->c_score = (int16_t)s; cl->c_k = (uint8_t)K; cl->c_off = (int8_t)off; cl->c_prob = pr; }
    __syncwarp();
    PROF(if (lane == 0) { sm->t_phase[5] += clock64() - t_y; sm->t_phase[6] += 1; sm->t_phase[7] += n_batch; })
    // Mate look-ahead: a candidate whose fewer end scored s will ask its mates for a score with limit <= K - s.
    // Each lane walks the mates of its own candidate and, per round, one still-unknown mate per lane is scored.
    {
        PROF(long long t_m = clock64();)
        bool walking = act_l && s >= 0;
        const int Km = K - (s > 0 ? s : 0);
        const uint32_t spl = act_l ? cl->set_pair : 0;
        const int dml = more == 0 ? (int)spl : 1 - (int)spl;
        const uint32_t cloc = act_l ? cl->loc : 0;
        uint32_t j = act_l ? cl->mate_index : 0;
        Mate *mbase = (mates + (size_t)spl * mate_cap);
        uint32_t n_done = 0;
        #pragma unroll 1
        for (int round = 0; round < MATE_LOOKAHEAD_ROUNDS; round++) {
            Mate *mt = nullptr;
            #pragma unroll 1
            while (walking) {
                Mate *q = &mbase[j];
                const bool needs = !is_within(q->loc, cloc, min_spacing) && q->best_possible <= (uint32_t)Km &&
                                   (q->score == (uint32_t)-2 || (q->score == (uint32_t)-1 && q->score_limit < (uint32_t)Km)) &&
                                   !mate_known(q, (uint32_t)Km);
                if (j == 0 || !is_within(mbase[j - 1].loc, cloc, max_spacing)) walking = false; else j--;
                if (needs) { mt = q; break; }
            }
            if (!__any_sync(FULL_MASK, mt != nullptr)) break;
            // the same mate may be wanted by several candidates of the batch: one lane scores it, with the largest limit
            const unsigned peers = __match_any_sync(FULL_MASK, (unsigned long long)mt);
            const int gmax = __reduce_max_sync(peers, Km);
            const bool mine = mt != nullptr && lane == __ffs((int)peers) - 1;
            int s2 = SC_NONE, off2 = 0;
            double pr2 = 0;
            score_location_lane(ix_slot, vm, dml, mine ? mt->loc : 0, mine ? mt->seed_offset : 0, gmax, lane_k, (lane_cell_t *)L + lane, lane_table + lane, mine,
                                &s2, &pr2, &off2);
            if (mine && s2 != SC_NONE) { mt->s_score = (int16_t)s2; mt->s_k = (uint8_t)gmax; mt->s_off = (int8_t)off2; mt->s_prob = pr2; }
            n_done += __popc(__ballot_sync(FULL_MASK, mine));
            __syncwarp();
        }
        PROF(if (lane == 0 && n_done) { sm->t_phase[5] += clock64() - t_m; sm->t_phase[6] += 1; sm->t_phase[7] += n_done; })
    }
}

// IntersectingPairedEndAligner::align.  All lanes.  v[0], v[1]: both mates staged (len set, Ns and non-ACGT bases
// counted by the caller: total_ns, n_bad[2]).
// Returns 0 = returned early leaving the result untouched, 1 = produced a result, 2 = scratch tier overflow.
__device__ int paired_intersect_warp(int ix_slot, const PairedCfg &cfg, const PairedScratch &sc, PairedSm *sm,
                                     uint8_t *rbase, uint32_t rlen0, uint32_t rlen1, uint32_t total_ns, uint32_t n_bad0, uint32_t n_bad1, uint8_t *W, int16_t *L,
                                     snapb200_paired_result *r, uint32_t pair_index, const MapqFixList &fix)
{
    const int lane = lane_id();
    const DevIndex &ix = c_index[ix_slot];
    // mate w staged at rbase + 4*w*rl (see ReadView); built on demand so that nothing is indexed dynamically
    auto view = [&](int w) { ReadView r = {rbase + (uint32_t)w * 4u * cfg.rl, cfg.rl, w ? rlen1 : rlen0}; return r; };
    const uint32_t seed_len = ix.seed_len, max_k = cfg.max_k, extra = cfg.extra;
    const uint32_t max_spacing = cfg.max_spacing, min_spacing = cfg.min_spacing;
    uint32_t max_seeds = cfg.num_seeds ? cfg.num_seeds : (uint32_t)(max(rlen0, rlen1) * cfg.seed_coverage / seed_len);
    if (max_seeds > MAX_LOOKUPS) max_seeds = MAX_LOOKUPS;  // rejected on the host; belt and braces
    if (rlen0 < 50 || rlen1 < 50) return 0;  // :186-188
    if (total_ns > max_k) return 0;              // :226-228

    PROF(long long t_a = clock64();)
    // ---- phase 1 (:259-340) ----
    if (lane == 0) {
        #pragma unroll 1
        for (int w = 0; w < 2; w++) {
            sm->popular[w] = 0; sm->n_look[w] = 0;
            #pragma unroll 1
            for (int d = 0; d < 2; d++) { sm->total_hits[w][d] = 0; sm->n_lookups[w][d] = 0; sm->cur_set[w][d] = -1; sm->inline_mask[w][d] = 0; }
        }
        sm->overflow = 0;
        sm->n_lv = 0;
        sm->n_probes = sm->n_hit_words = 0;
    }
    // Both mates' seeds are probed in one round (mate w on lanes 16w..) when they fit the 32 lanes, so the dependent
    // chain table entry -> overflow count word -> last hit word is paid once per pair, not once per mate.
    Phase1Sm *p1 = (Phase1Sm *)L;
    const bool together = max_seeds <= MAX_LOOKUPS / 2;
    #pragma unroll 1
    for (int pass = 0; pass < (together ? 1 : 2); pass++) {
        __syncwarp();
        if (lane == 0) {
            #pragma unroll 1
            for (int q = together ? 0 : pass; q <= (together ? 1 : pass); q++) {
                const uint32_t len_q = q ? rlen1 : rlen0, slot0 = together ? (uint32_t)q * (MAX_LOOKUPS / 2) : 0u;
                const bool plain = (q ? n_bad1 : n_bad0) == 0;  // no base that could invalidate a seed
                const uint32_t key = len_q | max_seeds << 16;
                const bool cacheable = plain && max_seeds <= MAX_LOOKUPS / 2;
                if (cacheable && sm->sc_key[q] == key) {
                    sm->n_sched_w[q] = sm->sc_n[q];
                    sm->sched_cached[q] = 1;
                } else {
                    schedule_seeds_paired(sm, p1, q, slot0, view(q).D(0), len_q, seed_len, max_seeds, plain);
                    sm->sched_cached[q] = 0;
                    if (cacheable) {
                        sm->sc_key[q] = key; sm->sc_n[q] = sm->n_sched_w[q];
                        #pragma unroll 1
                        for (uint32_t i = 0; i < sm->sc_n[q]; i++) { sm->sc_off[q][i] = p1->sched_off[slot0 + i]; sm->sc_wrap[q][i] = p1->sched_wrap[slot0 + i]; }
                    }
                }
            }
        }
        __syncwarp();
        const int w = together ? lane >> 4 : pass;
        const uint32_t jj = together ? (uint32_t)lane & 15u : (uint32_t)lane;
        if (jj < sm->n_sched_w[w]) {
            uint64_t sf, sr;
            HitList hl[2];
            uint32_t np = 0;
            pack_seed(view(w).D(0) + (sm->sched_cached[w] ? sm->sc_off[w][jj] : p1->sched_off[lane]), seed_len, &sf, &sr);
            lookup_seed(ix, sf, sr, hl, &np);
            #pragma unroll 1
            for (int d = 0; d < 2; d++) {
                p1->raw_hits[d][lane] = (unsigned long long)hl[d].hits;
                p1->raw_n[d][lane] = hl[d].n;
                p1->raw_last[d][lane] = (hl[d].n > 0 && hl[d].n < cfg.max_big_hits) ? __ldg(&hl[d].hits[hl[d].n - 1]) : 0xffffffffu;
            }
            atomicAdd(&sm->n_probes, np + (hl[0].n > 1) + (hl[1].n > 1));  // table slots + overflow count words
        }
        __syncwarp();
        // HashTableHitSet::recordLookup for every lookup of this round at once (:859-899).  In the reference this is a loop
        // over the lookups in schedule order; what it computes per direction -- which lookups start a new disjoint hit set
        // (the first one that is not too popular after each wrap of the seed schedule), the number of hit-less lookups per set,
        // the compacted list of lookups that have hits -- are prefix counts, i.e. ballots and popcounts.
        {
            const uint32_t slot = (uint32_t)lane;  // = jj, or 16 * mate + jj when both mates share the round
            const unsigned half = together ? 0xffffu << (16 * w) : FULL_MASK;
            const bool active = jj < sm->n_sched_w[w];
            const bool cached = sm->sched_cached[w] != 0;
            const uint32_t s_off = active ? (cached ? (uint32_t)sm->sc_off[w][jj] : (uint32_t)p1->sched_off[slot]) : 0u;
            const uint32_t s_wrap = active ? (cached ? (uint32_t)sm->sc_wrap[w][jj] : (uint32_t)p1->sched_wrap[slot]) : 0u;
            const uint32_t rlen_w = w ? rlen1 : rlen0;
            const unsigned le = (2u << lane) - 1u, lt = (1u << lane) - 1u;
            const bool writer = jj == 0;  // one lane per mate publishes the per-set totals
            #pragma unroll 1
            for (int d = 0; d < 2; d++) {
                uint32_t n = active ? p1->raw_n[d][slot] : 0u;
                const bool valid = active && n < cfg.max_big_hits;
                const unsigned popular_m = __ballot_sync(FULL_MASK, active && !valid) & half;
                // lookups of the same mate and wrap count that are not too popular: the lowest lane of each group starts a set
                const unsigned grp = __match_any_sync(FULL_MASK, valid ? (s_wrap | (uint32_t)w << 8) : (0x10000u | (uint32_t)lane));
                const unsigned starts = __ballot_sync(FULL_MASK, valid && lane == __ffs((int)grp) - 1) & half;
                const uint32_t set_id = (uint32_t)__popc(starts & le) - 1u;  // meaningful for valid lanes
                const bool empty = valid && n == 0, has = valid && n > 0;
                const unsigned has_m = __ballot_sync(FULL_MASK, has) & half;
                // sum of the hit counts of each mate (decides which mate has fewer hits, :342)
                const uint32_t sum0 = __reduce_add_sync(FULL_MASK, (valid && w == 0) ? n : 0u), sum1 = __reduce_add_sync(FULL_MASK, (valid && w == 1) ? n : 0u);
                if (jj < MAX_LOOKUPS / 2 || !together) sm->exhausted[w][d][jj] = 0;
                __syncwarp();
                const unsigned eg = __match_any_sync(FULL_MASK, empty ? (set_id | (uint32_t)w << 8) : (0x10000u | (uint32_t)lane));
                if (empty && lane == __ffs((int)eg) - 1) sm->exhausted[w][d][set_id] = (uint8_t)__popc(eg);
                const uint32_t offset = d == 0 ? s_off : rlen_w - seed_len - s_off;
                bool single = false;
                uint32_t k = 0;
                if (has) {
                    const uint32_t *hp = (const uint32_t *)p1->raw_hits[d][slot];
                    single = n == 1;  // the hit sits in the hash-table entry; raw_last is that hit
                    if (p1->raw_last[d][slot] < offset) {  // trim meaningless hits (:882-884); only at the very start of the genome
                        n--;
                        #pragma unroll 1
                        while (n > 0 && __ldg(&hp[n - 1]) < offset) n--;
                    }
                    k = (uint32_t)__popc(has_m & lt);
                    sm->hitref[w][d][k] = single ? p1->raw_last[d][slot] : (uint32_t)(hp - ix.overflow);
                    sm->nhits[w][d][k] = n;
                    sm->seedoff[w][d][k] = (uint16_t)offset;
                    sm->setid[w][d][k] = (uint8_t)set_id;
                }
                const uint32_t inl0 = __reduce_or_sync(FULL_MASK, (single && w == 0) ? 1u << k : 0u), inl1 = __reduce_or_sync(FULL_MASK, (single && w == 1) ? 1u << k : 0u);
                if (writer && (together || w == pass)) {
                    sm->total_hits[w][d] = w ? sum1 : sum0;
                    sm->cur_set[w][d] = (int8_t)(__popc(starts) - 1);
                    sm->n_lookups[w][d] = (uint8_t)__popc(has_m);
                    sm->inline_mask[w][d] = w ? inl1 : inl0;
                    sm->popular[w] += (uint32_t)__popc(popular_m);
                    sm->n_look[w] = sm->n_sched_w[w];
                }
                __syncwarp();
            }
        }
    }
    __syncwarp();
    const int more = (sm->total_hits[0][0] + sm->total_hits[0][1] > sm->total_hits[1][0] + sm->total_hits[1][1]) ? 0 : 1;  // :342
    const int fewer = 1 - more;
    // setPairDirection (:351): set pair sp uses read0 in direction sp, read1 in direction 1-sp
    if (lane == 0) {
        #pragma unroll 1
        sm->n_cands = 0; sm->n_mates[0] = sm->n_mates[1] = 0; sm->n_anchors = 0;
    }
    __syncwarp();

    PROF(long long t_b = clock64();)
    // ---- phase 2 (:359-511) ----
    uint32_t n_cands = 0;
    #pragma unroll 1
    for (int sp = 0; sp < 2; sp++) {
        const int d_fewer = fewer ? 1 - sp : sp, d_more = more ? 1 - sp : sp;  // set pair sp: read 0 in direction sp, read 1 in 1-sp
        LaneLookup lf = load_lookup(sm, ix.overflow, fewer, d_fewer);
        LaneLookup lm = load_lookup(sm, ix.overflow, more, d_more);
        const uint8_t *exh_f = sm->exhausted[fewer][d_fewer], *exh_m = sm->exhausted[more][d_more];
        const int cs_f = sm->cur_set[fewer][d_fewer], cs_m = sm->cur_set[more][d_more];
        uint32_t maxexh_f = 0, maxexh_m = 0;
        #pragma unroll 1
        for (int q = 0; q <= cs_f; q++) maxexh_f = max(maxexh_f, (uint32_t)exh_f[q]);
        #pragma unroll 1
        for (int q = 0; q <= cs_m; q++) maxexh_m = max(maxexh_m, (uint32_t)exh_m[q]);
        const uint32_t exhl_f = lf.act ? exh_f[lf.sid] : 0, exhl_m = lm.act ? exh_m[lm.sid] : 0;
        uint32_t mr_f = 0, mr_m = 0;  // mostRecentLocationReturned of each set
        uint32_t f_loc, f_off = 0, m_loc, m_off = 0;
        bool out_of_more = false;
        uint32_t n_mates = 0, last_mate_loc = 0;
        Mate *mates = sc.mates_of(sp);
        if (!hs_first(lf, &mr_f, &f_loc, &f_off)) continue;
        m_loc = INVALID_LOC;
        #pragma unroll 1
        for (;;) {
            if (m_loc > f_loc + max_spacing) {
                if (!hs_next_le(lm, &mr_m, f_loc + max_spacing, &m_loc, &m_off)) break;
            }
            if (m_loc + max_spacing < f_loc && (n_mates == 0 || !is_within(last_mate_loc, f_loc, max_spacing))) {
                if (!hs_next_le(lf, &mr_f, m_loc + max_spacing, &f_loc, &f_off)) break;
                continue;
            }
            #pragma unroll 1
            while (m_loc + max_spacing >= f_loc && !out_of_more) {
                uint32_t bp = hs_best_possible(lm, exhl_m, maxexh_m, mr_m, max_k);
                if (n_mates >= cfg.mate_cap) return 2;
                if (lane == 0) {
                    Mate *m = &mates[n_mates];
                    m->loc = m_loc; m->best_possible = bp; m->seed_offset = m_off;
                    m->score = (uint32_t)-2; m->score_limit = (uint32_t)-1; m->prob = 0; m->genome_offset = 0;
                    m->s_score = SC_NONE; m->s_k = 0;
                    sm->ring_loc[n_mates & 31] = m_loc; sm->ring_bp[n_mates & 31] = bp;
                }
                n_mates++;
                last_mate_loc = m_loc;
                if (!hs_next_lower(lm, &mr_m, &m_loc, &m_off)) {
                    m_loc = 0;
                    out_of_more = true;
                    break;
                }
            }
            uint32_t bp_fewer = hs_best_possible(lf, exhl_f, maxexh_f, mr_f, max_k);
            // lowest bestPossibleScore among the mates in range (:469-475): scan back from the newest mate, 32 per step
            uint32_t low_mate = max_k + extra;
            __syncwarp();  // the mate records and the ring must be visible to all lanes
            #pragma unroll 1
            for (int top = (int)n_mates - 1; top >= 0; top -= 32) {
                const int i = top - lane;
                bool stop = false;
                uint32_t bp = 0xffffffffu;
                if (i >= 0) {
                    // the newest 32 mates are in the shared-memory ring; older ones (rare: more than 32 mates in range) in HBM
                    const bool recent = top == (int)n_mates - 1;
                    const uint32_t mloc = recent ? sm->ring_loc[i & 31] : mates[i].loc;
                    if (mloc > f_loc + max_spacing) stop = true; else bp = recent ? sm->ring_bp[i & 31] : mates[i].best_possible;
                }
                const unsigned stops = __ballot_sync(FULL_MASK, stop);
                if (stops) {  // lanes beyond the first out-of-range mate do not count
                    const int first = __ffs((int)stops) - 1;
                    if (lane >= first) bp = 0xffffffffu;
                }
                low_mate = min(low_mate, __reduce_min_sync(FULL_MASK, bp));
                if (stops) break;
            }
            if (low_mate + bp_fewer <= max_k + extra) {
                if (n_cands >= cfg.cand_cap) return 2;
                if (lane == 0) {
                    Cand *c = &sc.cands[n_cands];
                    c->loc = f_loc; c->set_pair = (uint8_t)sp; c->mate_index = n_mates - 1; c->seed_offset = (uint16_t)f_off;
                    c->best_possible = (uint8_t)bp_fewer; c->list = (int)(low_mate + bp_fewer); c->anchor = -1;
                    c->c_score = SC_NONE; c->c_k = 0;
                }
                n_cands++;
            }
            if (!hs_next_lower(lf, &mr_f, &f_loc, &f_off)) break;
        }
        uint32_t words = __reduce_add_sync(FULL_MASK, lf.words + lm.words);
        if (lane == 0) sm->n_hit_words += words;
    }
    __syncwarp();

    PROF(long long t_c = clock64(); long long t_lv = 0;)
    // ---- phase 3 (:516-720) ----
    // The visiting order of candidates (lists 0,1,2..., LIFO inside a list) and of a candidate's mates is fixed once
    // phase 2 is done, and an LV result for limit k is (d <= k ? (d, probability, netIndel) : -1) with d, probability and
    // netIndel independent of k.  So scores are computed AHEAD of their use, 32 locations per warp (one per lane,
    // lv_lane) with the current limit, which only ever shrinks, and are committed one by one in the reference's order
    // with the limit in force at that moment.  Counters (nLocationsScored) advance at commit time only.
    // The reference keeps one LIFO list of candidates per bestPossibleScore and serves list 0, 1, 2, ... (:516-530).  Every
    // list is filled in ascending candidate index, so the visiting order is: by list, then by descending index -- a
    // counting sort, done here by the whole warp, instead of 32 linked lists that the leader would have to chase through HBM.
    {
        sm->list_pos[lane] = 0;
        __syncwarp();
        #pragma unroll 1
        for (uint32_t b0 = 0; b0 < n_cands; b0 += 32) {
            const uint32_t i = b0 + lane;
            const uint32_t l = i < n_cands ? (uint32_t)sc.cands[i].list : 0xffffffffu;
            const unsigned peers = __match_any_sync(FULL_MASK, l);
            if (i < n_cands && lane == __ffs((int)peers) - 1) sm->list_pos[l] += __popc(peers);
            __syncwarp();
        }
        uint32_t cnt = sm->list_pos[lane], start = cnt;  // inclusive scan over the (at most 31) lists
        #pragma unroll 1
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(FULL_MASK, start, o);
            if (lane >= o) start += up;
        }
        __syncwarp();
        sm->list_pos[lane] = start - cnt;
        __syncwarp();
        #pragma unroll 1
        for (int top = (int)n_cands - 1; top >= 0; top -= 32) {
            const int i = top - lane;  // lane 0 holds the highest index: visited first within its list
            const uint32_t l = i >= 0 ? (uint32_t)sc.cands[i].list : 0xffffffffu;
            const unsigned peers = __match_any_sync(FULL_MASK, l);
            if (i >= 0) sc.order[sm->list_pos[l] + __popc(peers & ((1u << lane) - 1u))] = (uint32_t)i;
            __syncwarp();
            if (i >= 0 && lane == __ffs((int)peers) - 1) sm->list_pos[l] += __popc(peers);
            __syncwarp();
        }
    }
    // Phase 3 proper is a leader-driven state machine.  Lane 0 walks candidates and their mates in the reference's order for
    // as long as every score it needs is already known (from the look-ahead below, or because the location was scored
    // before): candidates whose fewer end does not fit the limit, mates that are out of range or already failed, commits
    // and merges cost no warp-wide step at all.  Only when a score is missing does it hand over to the warp, which scores
    // that location together with the next ones in visiting order, and then picks up exactly where it stopped.
    enum { NEED_DONE = 0, NEED_OVERFLOW, NEED_CAND_BATCH, NEED_CAND_WARP, NEED_MATE };
    enum { ST_PICK = 0, ST_PICK_RESUME, ST_MATES, ST_MATES_RESUME };
    if (lane == 0) {
        sm->n_cands = n_cands;
        sm->pos = 0;
        sm->score_limit = max_k + extra;
        sm->p_all = 0; sm->p_best = 0;
        sm->best_pair_score = 65536;
        sm->stop = 0;
        sm->state = ST_PICK;
    }
    __syncwarp();
    #pragma unroll 1
    for (;;) {
        if (lane == 0) {
            int need = -1;
            #pragma unroll 1
            while (need < 0) {
                if (sm->state <= ST_PICK_RESUME) {  // the next candidate in visiting order (:516-557)
                    const bool resume = sm->state == ST_PICK_RESUME;
                    if (!resume && (sm->stop || sm->pos >= n_cands)) { need = NEED_DONE; break; }
                    const int ci = resume ? sm->ci : (int)sc.order[sm->pos];
                    const Cand *c = &sc.cands[ci];
#ifdef PF_CANDS
                    // the leader's loads are a dependent chain through HBM-resident scratch (order -> candidate -> its mates): ask
                    // for the records of the candidates a few positions ahead now, so that they are in L1 when their turn comes
                    if (!resume) {
                        if (sm->pos + PF_CANDS < n_cands) {
                            const uint32_t nj = sc.order[sm->pos + PF_CANDS];
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(&sc.cands[nj]));
                        }
                        if (sm->pos + PF_CANDS / 2 < n_cands) {
                            const Cand *nc = &sc.cands[sc.order[sm->pos + PF_CANDS / 2]];
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(&sc.mates_of(nc->set_pair)[nc->mate_index]));
                        }
                    }
#endif
                    if (!resume) {
                        if ((uint32_t)c->list > sm->score_limit) { need = NEED_DONE; break; }  // lists beyond the limit are never served (:527-530)
                        sm->n_lv++;
                    }
                    const int cs = c->c_score;
                    if (cs == SC_NONE) {  // not scored yet: the warp scores it and the next unscored candidates in visiting order
                        sm->ci = ci; sm->c_loc = c->loc; sm->c_seedoff = c->seed_offset; sm->c_sp = c->set_pair;
                        sm->state = ST_PICK_RESUME;
                        need = sm->score_limit < cfg.lane_gate ? NEED_CAND_BATCH : NEED_CAND_WARP;
                        break;
                    }
                    if (!(cs >= 0 && (uint32_t)cs <= sm->score_limit)) {  // scoreLocation says -1 for the limit in force: next candidate
                        sm->pos++;
                        sm->state = ST_PICK;
                        continue;
                    }
                    sm->ci = ci; sm->c_loc = c->loc; sm->c_sp = c->set_pair; sm->mi = c->mate_index;
                    sm->fs = cs; sm->f_prob = c->c_prob; sm->f_off = (int)c->c_off;
                    sm->state = ST_MATES;
                }
                // mates of the current candidate (:559-711)
                const uint32_t sp = sm->c_sp;
                const int dir_f = fewer == 0 ? (int)sp : 1 - (int)sp, dir_m = more == 0 ? (int)sp : 1 - (int)sp;
                const uint32_t f_score = (uint32_t)sm->fs;
                const double f_prob = sm->f_prob;
                const int f_off = sm->f_off;
                Mate *mbase = sc.mates_of(sp);
                Cand *c = &sc.cands[sm->ci];
                #pragma unroll 1
                for (;;) {
                    Mate *m = &mbase[sm->mi];
#ifdef PF_MATES
                    if (sm->mi >= PF_MATES) asm volatile("prefetch.global.L1 [%0];" ::"l"(&mbase[sm->mi - PF_MATES]));  // the walk goes down the array
#endif
                    int act = 0;
                    if (sm->state == ST_MATES_RESUME) {  // the warp has just scored this mate for sm->m_limit
                        act = 2;
                        sm->state = ST_MATES;
                    } else if (!is_within(m->loc, sm->c_loc, min_spacing) && m->best_possible <= sm->score_limit - f_score) {
                        act = 1;
                        const uint32_t m_limit = sm->score_limit - f_score;
                        if (m->score == (uint32_t)-2 || (m->score == (uint32_t)-1 && m->score_limit < m_limit)) {
                            act = 2;
                            sm->m_limit = m_limit;
                            sm->n_lv++;
                            if (!mate_known(m, m_limit)) {
                                // not known well enough: gather the mates further down this candidate's range that will need a score
                                sm->m_loc = m->loc; sm->m_seedoff = m->seed_offset;
                                uint32_t nb = 0;
                                const bool lane_ok = m_limit < cfg.lane_gate;
                                uint32_t j = sm->mi;
                                #pragma unroll 1
                                for (;;) {
                                    Mate *q = &mbase[j];
                                    if (!is_within(q->loc, sm->c_loc, min_spacing) && q->best_possible <= m_limit &&
                                        (q->score == (uint32_t)-2 || (q->score == (uint32_t)-1 && q->score_limit < m_limit)) &&
                                        !mate_known(q, m_limit))
                                        sm->batch_ids[nb++] = j;
                                    if (!lane_ok || nb >= 32) break;
                                    if (j == 0 || !is_within(mbase[j - 1].loc, sm->c_loc, max_spacing)) break;
                                    j--;
                                }
                                sm->n_batch = nb;
                                sm->state = ST_MATES_RESUME;
                                need = NEED_MATE;
                                break;
                            }
                        }
                    }
                    if (act == 2) {  // commit: what scoreLocation(limit = m_limit) returns
                        const int d = m->s_score;
                        const bool ok = d >= 0 && (uint32_t)d <= sm->m_limit;
                        m->score = ok ? (uint32_t)d : (uint32_t)-1;
                        m->prob = ok ? m->s_prob : 0.0;
                        m->genome_offset = ok ? (int)m->s_off : 0;
                        m->score_limit = sm->m_limit;
                    }
                    if (act != 0 && m->score != (uint32_t)-1) {
                        const double pair_prob = m->prob * f_prob;
                        const uint32_t pair_score = m->score + f_score;
                        const uint32_t new_more = m->loc + (uint32_t)m->genome_offset, new_fewer = c->loc + (uint32_t)f_off;
                        int an = c->anchor;
                        const int ci = sm->ci;
                        if (an < 0) {  // look for a merge anchor among neighbouring candidates (:598-627)
                            #pragma unroll 1
                            for (int j = ci - 1; j >= 0 && is_within(sc.cands[j].loc, new_fewer, 50) && sc.cands[j].set_pair == c->set_pair; j--) {
                                if (sc.cands[j].anchor >= 0) { c->anchor = an = sc.cands[j].anchor; break; }
                            }
                            if (an < 0) {
                                // the reference's second scan starts one above and walks DOWN (:615-619); below index 0 it
                                // reads out of bounds there, which is treated as the end of the scan here
                                #pragma unroll 1
                                for (int j = ci + 1; j >= 0 && j < (int)sm->n_cands && is_within(sc.cands[j].loc, new_fewer, 50) &&
                                                 sc.cands[j].set_pair == c->set_pair; j--) {
                                    if (sc.cands[j].anchor >= 0) { c->anchor = an = sc.cands[j].anchor; break; }
                                }
                            }
                        }
                        bool merged;
                        double old_prob;
                        if (an < 0) {
                            if (sm->n_anchors >= cfg.anchor_cap) {
                                sm->overflow = 1;
                                merged = true;
                                old_prob = 0;
                            } else {
                                Anchor *ma = &sc.anchors[sm->n_anchors];
                                ma->loc_more = new_more; ma->loc_fewer = new_fewer; ma->prob = pair_prob; ma->pair_score = (int)pair_score;
                                c->anchor = (int)sm->n_anchors++;
                                merged = false;
                                old_prob = 0;
                            }
                        } else {  // MergeAnchor::checkMerge (:1324-1371)
                            Anchor *ma = &sc.anchors[an];
                            uint32_t dm = ma->loc_more > new_more ? ma->loc_more - new_more : new_more - ma->loc_more;
                            uint32_t df = ma->loc_fewer > new_fewer ? ma->loc_fewer - new_fewer : new_fewer - ma->loc_fewer;
                            if (ma->loc_more == INVALID_LOC || !(dm < 50 && df < 50)) {
                                ma->loc_more = new_more; ma->loc_fewer = new_fewer; ma->prob = pair_prob; ma->pair_score = (int)pair_score;
                                old_prob = 0;
                                merged = false;
                            } else if ((int)pair_score < ma->pair_score || ((int)pair_score == ma->pair_score && pair_prob > ma->prob)) {
                                old_prob = ma->prob;
                                ma->prob = pair_prob;
                                ma->pair_score = (int)pair_score;
                                merged = false;
                            } else {
                                old_prob = 0;
                                merged = true;
                            }
                        }
                        if (!merged) {
                            double t = sm->p_all - old_prob;
                            sm->p_all = 0 > t ? 0 : t;
                            if (pair_score <= max_k && (pair_score < sm->best_pair_score ||
                                                        (pair_score == sm->best_pair_score && pair_prob > sm->p_best))) {
                                sm->best_pair_score = pair_score;
                                sm->p_best = pair_prob;
                                sm->best_loc[fewer] = new_fewer; sm->best_loc[more] = new_more;
                                sm->best_score[fewer] = f_score; sm->best_score[more] = m->score;
                                sm->best_dir[fewer] = dir_f; sm->best_dir[more] = dir_m;
                                sm->score_limit = pair_score + extra;
                            }
                            sm->p_all += pair_prob;
                            if (sm->p_all >= 4.9) sm->stop = 1;  // nothing rescues a 0 MAPQ (:693-698)
                        }
                    }
                    bool go_on = true;
                    if (sm->stop || sm->overflow) go_on = false;
                    else if (sm->mi == 0 || !is_within(mbase[sm->mi - 1].loc, c->loc, max_spacing)) go_on = false;
                    else sm->mi--;
                    if (!go_on) {  // done with this candidate
                        if (sm->overflow) need = NEED_OVERFLOW;
                        else { if (!sm->stop) sm->pos++; sm->state = ST_PICK; }
                        break;
                    }
                }
            }
            sm->act = need;
        }
        __syncwarp();
        const int need = sm->act;
        if (need == NEED_DONE) break;
        if (need == NEED_OVERFLOW) return 2;
        PROF(long long t_x = clock64();)
        const uint32_t sp = sm->c_sp;
        const int dir_f = fewer == 0 ? (int)sp : 1 - (int)sp, dir_m = more == 0 ? (int)sp : 1 - (int)sp;
        if (need == NEED_MATE) {
            const uint32_t nb = sm->n_batch;
            const int K = (int)sm->m_limit;
            if (nb >= LANE_MIN_BATCH) {
                PROF(long long t_y = clock64();)
                bool act_l = (uint32_t)lane < nb;
                Mate *ml = act_l ? &sc.mates_of(sp)[sm->batch_ids[lane]] : nullptr;
                int s = SC_NONE, off = 0;
                double pr = 0;
                score_location_lane(ix_slot, view(more), dir_m, act_l ? ml->loc : 0, act_l ? ml->seed_offset : 0, K, (int)cfg.lane_k, (lane_cell_t *)L + lane, sc.lane_table + lane, act_l, &s, &pr, &off);
                if (act_l && s != SC_NONE) { ml->s_score = (int16_t)s; ml->s_k = (uint8_t)K; ml->s_off = (int8_t)off; ml->s_prob = pr; }
                __syncwarp();
                PROF(if (lane == 0) { sm->t_phase[5] += clock64() - t_y; sm->t_phase[6] += 1; sm->t_phase[7] += nb; })
            }
            if (lane == 0) { const Mate *m = &sc.mates_of(sp)[sm->mi]; sm->act2 = !mate_known(m, sm->m_limit); }
            __syncwarp();
            if (sm->act2) {
                double m_prob;
                int m_off;
                PROF(long long t_w = clock64();)
                int ms = score_location_warp(ix_slot, view(more), dir_m, sm->m_loc, sm->m_seedoff, K, false, W, L, &m_prob, &m_off);
                __syncwarp();
                PROF(if (lane == 0) { sm->t_phase[8] += clock64() - t_w; sm->t_phase[9] += 1; })
                if (lane == 0) { Mate *m = &sc.mates_of(sp)[sm->mi]; m->s_score = (int16_t)ms; m->s_k = (uint8_t)K; m->s_off = (int8_t)m_off; m->s_prob = m_prob; }
            }
            __syncwarp();
            PROF(t_lv += clock64() - t_x;)
            continue;
        }
        // a candidate's fewer end
        uint32_t n_batch = 0;
        if (need == NEED_CAND_BATCH) {  // the unscored candidates among the next 32 positions of the visiting order
            const uint32_t at = sm->pos + (uint32_t)lane;
            bool want = false;
            uint32_t j = 0;
            if (at < n_cands) {
                j = sc.order[at];
                want = sc.cands[j].c_score == SC_NONE && (uint32_t)sc.cands[j].list <= sm->score_limit;
            }
            const unsigned m = __ballot_sync(FULL_MASK, want);
            if (want) sm->batch_ids[__popc(m & ((1u << lane) - 1u))] = j;
            n_batch = (uint32_t)__popc(m);
            __syncwarp();
        }
        if (n_batch >= LANE_MIN_BATCH)
            lane_batch_candidates(ix_slot, (int)cfg.lane_k, min_spacing, max_spacing, sc.cands, sc.mates, sc.mate_cap, sc.lane_table, sm, view(fewer), view(more), fewer, n_batch, L);
        if (lane == 0) sm->act2 = sc.cands[sm->ci].c_score == SC_NONE;
        __syncwarp();
        if (sm->act2) {  // warp mode for this one candidate (small batch, large limit, or a window at the genome's edge)
            double pr;
            int off;
            const int K = (int)sm->score_limit;
            PROF(long long t_w = clock64();)
            int s = score_location_warp(ix_slot, view(fewer), dir_f, sm->c_loc, sm->c_seedoff, K, false, W, L, &pr, &off);
            __syncwarp();
            PROF(if (lane == 0) { sm->t_phase[8] += clock64() - t_w; sm->t_phase[9] += 1; })
            if (lane == 0) { Cand *c = &sc.cands[sm->ci]; c->c_score = (int16_t)s; c->c_k = (uint8_t)K; c->c_off = (int8_t)off; c->c_prob = pr; }
        }
        __syncwarp();
        PROF(t_lv += clock64() - t_x;)
    }

    if (lane == 0) {
        PROF(long long t_d = clock64();)
        PROF(sm->t_phase[1] = t_b - t_a; sm->t_phase[2] = t_c - t_b; sm->t_phase[3] = t_lv; sm->t_phase[4] = (t_d - t_c) - t_lv;)
        if (sm->best_pair_score == 65536) {
            #pragma unroll 1
            for (int w = 0; w < 2; w++) {
                r->location[w] = INVALID_LOC; r->mapq[w] = 0; r->score[w] = -1; r->status[w] = SNAPB200_NOT_FOUND;
            }
        } else {
            #pragma unroll 1
            for (int w = 0; w < 2; w++) {
                bool near_int;
                int popular = (int)(sm->popular[0] + sm->popular[1]);
                int mq = compute_mapq_dev(sm->p_all, sm->p_best, (int)sm->best_score[w], popular, &near_int);
                if (near_int) {
                    uint32_t slot = atomicAdd(fix.count, 1u);
                    if (slot < fix.cap) {
                        MapqFix f;
                        f.index = pair_index; f.end = (uint32_t)w; f.p_all = sm->p_all; f.p_best = sm->p_best;
                        f.score = (int)sm->best_score[w]; f.popular = popular; f.divisor = 1; f.is_paired_rule = 1;
                        fix.items[slot] = f;
                    }
                }
                r->location[w] = sm->best_loc[w];
                r->direction[w] = (uint8_t)sm->best_dir[w];
                r->mapq[w] = mq;
                r->status[w] = mq > 10 ? SNAPB200_SINGLE_HIT : SNAPB200_MULTIPLE_HITS;
                r->score[w] = (int)sm->best_score[w];
            }
        }
        r->p_all = sm->p_all;
        r->p_best = sm->p_best;
    }
    __syncwarp();
    return 1;
}
