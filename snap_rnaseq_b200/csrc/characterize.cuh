// characterize.cuh -- BaseAligner::CharacterizeSeeds for a batch, one warp per read.
//
// Replaces BaseAligner::CharacterizeSeeds (SNAPLib/BaseAligner.cpp:206-508; callers AlignmentFilter.cpp:758, 968-971):
// stages 1-2 of the single-end aligner without any scoring -- the seed schedule of AlignRead, one lookupSeed per
// seed, and for every hit of a direction that is not too popular the tuple (genome location of the read start,
// forward seed offset) inserted into that direction's std::map<location, std::set<seedOffset>>.
//
// Device form of the two maps: segment s = 2*read + direction holds that map's (location, seedOffset) tuples in
// ascending order, which is the in-order traversal of the reference's containers.  The kernel runs twice: COUNT
// sizes the segments, EMIT writes one 64-bit key (segment | location | seedOffset) per tuple; a radix sort over
// the key bits that vary puts every segment in order, and a last pass splits the keys into the output arrays.
// The seeds of a read are probed up to 32 at a time, one per lane, as in single.cuh; hit lists are read 32 words
// per step (coalesced) and the slot of each tuple comes from a ballot + popcount, so no atomics are needed.
#pragma once
#include "single.cuh"

#define CHAR_OFF_BITS 9   // seed offsets < MAX_READ_LENGTH (500) fit in 9 bits
#define CHAR_LOC_BITS 32
#define CHAR_SEG_SHIFT (CHAR_OFF_BITS + CHAR_LOC_BITS)

struct CharArgs {
    DevIndex ix;
    DevBatch b;
    uint32_t max_hits, max_k, num_seeds, explore;
    double seed_coverage;
    uint32_t rl;                        // shared-memory bytes reserved per read
    unsigned long long *seg;            // COUNT: [2n] tuple counts (out); EMIT: [2n] exclusive offsets (in)
    unsigned long long *keys;           // EMIT
    Counters *ctr;
};

__host__ __device__ inline size_t char_warp_shared(uint32_t rl)
{
    return ((sizeof(SingleSm) + 15) & ~(size_t)15) + (((size_t)rl + 15) & ~(size_t)15);
}

template <bool EMIT>
__global__ void __launch_bounds__(256) characterize_kernel(const CharArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    uint8_t *base = smem + char_warp_shared(a.rl) * warp;
    SingleSm *sm = (SingleSm *)base;
    uint8_t *D = base + ((sizeof(SingleSm) + 15) & ~(size_t)15);
    const uint32_t seed_len = a.ix.seed_len;
    for (;;) {
        const uint32_t r = fetch_work(&a.ctr->work);
        if (r >= a.b.n) break;
        const uint32_t off = a.b.offsets[r], len = a.b.offsets[r + 1] - off;
        unsigned long long cnt[2] = {0, 0};
        unsigned long long seg_base[2] = {0, 0};
        if (EMIT) { seg_base[0] = a.seg[2 * r]; seg_base[1] = a.seg[2 * r + 1]; }
        bool go = len >= seed_len;  // "Too short to have any seeds" (:277-282)
        if (go) {
            uint32_t ns = 0;
            for (uint32_t i0 = 0; i0 < len; i0 += 32) {
                const uint32_t i = i0 + lane;
                uint8_t c = 0;
                if (i < len) { c = a.b.bases[off + i]; D[i] = c; }
                ns += __popc(__ballot_sync(FULL_MASK, c == 'N'));
            }
            go = ns <= a.max_k;  // :303-306
        }
        __syncwarp();
        if (go) {
            const uint32_t max_seeds = a.num_seeds ? a.num_seeds : (uint32_t)(int)(a.seed_coverage * len / seed_len);
            if (lane == 0) {
                for (int i = 0; i < 16; i++) sm->used[i] = 0;
                sm->next = 0; sm->wrap = 0;
            }
            uint32_t applied = 0;  // nSeedsApplied[FORWARD] + nSeedsApplied[RC]; uniform across the warp
            for (;;) {
                if (applied >= max_seeds) break;
                __syncwarp();
                if (lane == 0) schedule_seeds_single(sm, D, len, seed_len);
                __syncwarp();
                const uint32_t n_sched = sm->n_sched;
                HitList my[2] = {{nullptr, 0}, {nullptr, 0}};
                if ((uint32_t)lane < n_sched) {
                    uint64_t sf, sr;
                    pack_seed(D + sm->sched_off[lane], seed_len, &sf, &sr);
                    lookup_seed(a.ix, sf, sr, my, nullptr);
                }
                bool out = false;
                for (uint32_t j = 0; j < n_sched; j++) {
                    if (applied >= max_seeds) { out = true; break; }
                    const uint32_t seed_at = sm->sched_off[j];
                    for (int dir = 0; dir < 2; dir++) {
                        const uint32_t n = __shfl_sync(FULL_MASK, my[dir].n, (int)j);
                        const uint32_t *hits = (const uint32_t *)shfl_u64((uint64_t)my[dir].hits, (int)j);
                        if (n > a.max_hits && !a.explore) continue;  // popular seed: pretend we never looked (:394-401)
                        const uint32_t offset = dir == 0 ? seed_at : len - seed_len - seed_at;
                        const uint32_t lim = min(n, a.max_hits);
                        for (uint32_t b0 = 0; b0 < lim; b0 += 32) {
                            const uint32_t i = b0 + lane;
                            uint32_t hit = 0;
                            bool ok = false;
                            if (i < lim) { hit = __ldg(&hits[i]); ok = hit >= offset; }  // :446-450 with the full location range
                            const unsigned m = __ballot_sync(FULL_MASK, ok);
                            if (EMIT && ok) {
                                const unsigned long long pos = seg_base[dir] + cnt[dir] + __popc(m & ((1u << lane) - 1u));
                                // the forward seed offset goes into BOTH maps (:455-478)
                                a.keys[pos] = ((unsigned long long)(2 * r + dir) << CHAR_SEG_SHIFT) |
                                              ((unsigned long long)(hit - offset) << CHAR_OFF_BITS) | seed_at;
                            }
                            cnt[dir] += __popc(m);
                        }
                        applied++;
                    }
                }
                if (out || sm->terminal) break;  // terminal: wrapCount >= seedLen (:330-337)
            }
        }
        if (!EMIT && lane == 0) { a.seg[2 * r] = cnt[0]; a.seg[2 * r + 1] = cnt[1]; }
        __syncwarp();
    }
}

__global__ void characterize_split_kernel(const unsigned long long *keys, unsigned long long n, uint32_t *locs, uint16_t *offs)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = keys[i];
    locs[i] = (uint32_t)(k >> CHAR_OFF_BITS);
    offs[i] = (uint16_t)(k & ((1u << CHAR_OFF_BITS) - 1u));
}
