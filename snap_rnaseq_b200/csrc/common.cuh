// common.cuh -- shared declarations of the B200 alignment core (device structs, error handling).
//
// Execution model used by every kernel in this library:
//   * one warp aligns one read (or one pair); a CTA is WARPS_PER_CTA independent warps; the grid is
//     persistent (a multiple of the SM count) and warps pull work items from a global atomic counter;
//   * the reference algorithm is sequential and order dependent, so per-read
//     bookkeeping runs in lane 0 ("leader sections") while the data-parallel parts -- seed packing,
//     hash-table probes, hit-list loads, candidate-bucket probes, Landau-Vishkin diagonals, genome window
//     staging -- run on all 32 lanes ("warp sections");
//   * everything that crosses between the two lives in a per-warp block of shared memory and is read only
//     after a __syncwarp(), so all lanes follow the same control flow.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/snapb200.h"

// Cycle accounting of paired_kernel (scripts/profrun.py) is compiled in only with -DSNAPB200_PROFILE: the kernels are
// instruction-cache bound, so code that is not needed is kept out of them.
#ifdef SNAPB200_PROFILE
#define PROF(...) __VA_ARGS__
#else
#define PROF(...)
#endif

#define FULL_MASK 0xffffffffu
#define MAXK SNAPB200_MAX_K           // 31
#define INVALID_LOC 0xffffffffu
#define UNUSED_SCORE 0xffffu
#define BUCKET 48u                    // BaseAligner::maxMergeDist == hashTableElementSize (BaseAligner.h:163,196)
#define GENOME_PAD 128                // 'n' bytes kept before and after the genome in HBM (reference: 100, Genome.h:175)
#define WIN_SLACK 40                  // genome bytes staged on each side of [loc, loc+readLen)
#define MAX_LOOKUPS 32                // lookups held per hit set (one per lane)
#define STATUS_RETRY 0xfe             // internal: scratch tier too small, rerun in the large tier

struct HtEntry { uint32_t key, v1, v2; };  // SNAPHashTable::Entry, HashTable.h:119-123

// Index + genome + probability tables resident in HBM.
struct DevIndex {
    const HtEntry *tables;         // all hash tables back to back
    const uint64_t *table_start;   // [n_tables] first entry of each table
    const uint64_t *table_size;    // [n_tables]
    const uint32_t *overflow;      // OverflowTable words
    const uint8_t *genome;         // byte 0 of the genome; GENOME_PAD readable 'n' bytes on both sides
    const uint32_t *piece_begin;   // [n_pieces]
    const double *phred;           // lv_phredToProbability[256]
    const double *indel;           // lv_indelProbabilities[64]
    const double *perfect;         // lv_perfectMatchProbability[501]
    double seed_prob;              // pow(1-SNP_PROB, seedLen) as the reference's build evaluates it
    uint32_t n_bases, n_pieces, seed_len, n_tables, padding;
};

// The descriptors of the indices open on this device.  Kernels receive a slot number: out-of-line device functions then
// read the descriptor through the constant cache instead of through a by-reference copy on the caller's stack (which was
// most of the kernels' local-memory traffic).
#define MAX_INDEX_SLOTS 16
__constant__ DevIndex c_index[MAX_INDEX_SLOTS];  // the library is one translation unit (snapb200.cu)

// mapq fix-up request: the device's log10 landed within 1e-9 of an integer, where a last-ulp difference
// from glibc could change the truncation; the host re-evaluates computeMAPQ for these with libm.
struct MapqFix {
    uint32_t index;    // read / pair index in the batch
    uint32_t end;      // 0/1 for pairs
    double p_all, p_best;
    int32_t score, popular, divisor, is_paired_rule;  // is_paired_rule: status uses mapq > 10 instead of >= 10
};

extern thread_local char g_last_error[512];
int set_error(int code, const char *fmt, ...);

#define CUDA_TRY(expr)                                                                           \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return set_error(SNAPB200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                                \
    } while (0)

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// broadcast helpers
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src)
{
    uint32_t lo = __shfl_sync(FULL_MASK, (uint32_t)v, src);
    uint32_t hi = __shfl_sync(FULL_MASK, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ double shfl_f64(double v, int src)
{
    return __longlong_as_double((long long)shfl_u64((uint64_t)__double_as_longlong(v), src));
}
