// index_build.cuh -- sort-based genome-index construction on the device.
//
// Produces an index that is lookup-equivalent to the one GenomeIndex::BuildIndexToDirectory builds
// (SNAPLib/GenomeIndex.cpp:348-720, worker :1394-1496, ApplyHashTableUpdate :1498-1579): for every seed,
// lookupSeed returns the same (nHits, hit list) in both directions -- hit lists are sorted descending
// (GenomeIndex.cpp:566-619), singletons are stored inline, multi-hit seeds point into the overflow table
// (count word followed by the locations) -- and the open-addressed tables use the same hash and probe
// sequence (HashTable.h:60-105), so the reference can load and use a saved copy.  Slot order inside a table
// differs from a reference-built index, which already depends on its thread interleaving.
//
// Instead of the reference's locked inserts + per-seed linked lists, the device version is
//   1. one thread per genome position: 2-bit pack the seed and its reverse complement, keep the smaller
//      ("canonical") one plus a which-strand bit;
//   2. one stable radix sort of (canonical<<1|strand, position) with positions fed in descending order, so
//      each hit list comes out contiguous and already sorted descending;
//   3. run detection + two prefix sums to lay out the overflow table;
//   4. lock-free insertion: keys are unique, so claiming the first empty slot of the probe sequence with an
//      atomicCAS on value1 is enough.
#pragma once
#include <cub/cub.cuh>

#include "common.cuh"
#include "lookup.cuh"

#define IB_INVALID_KEY 0xffffffffffffffffull

// record i describes genome position (n_pos-1-i): descending positions + stable sort = descending hit lists
__global__ void ib_emit_kernel(const uint8_t *genome, uint32_t n_pos, uint32_t seed_len, unsigned long long *keys, uint32_t *vals,
                               unsigned long long *n_valid)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = false;
    if (i < n_pos) {
        uint32_t pos = n_pos - 1 - i;
        uint64_t f, r;
        valid = pack_seed(genome + pos, seed_len, &f, &r);
        uint64_t canon = f > r ? r : f;
        uint64_t strand = f > r ? 1 : 0;  // 1: stored under the reverse complement's entry (usingComplement)
        keys[i] = valid ? ((canon << 1) | strand) : IB_INVALID_KEY;
        vals[i] = pos;
    }
    unsigned b = __ballot_sync(FULL_MASK, valid);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(n_valid, (unsigned long long)__popc(b));
}

__global__ void ib_heads_kernel(const unsigned long long *keys, uint32_t n, uint32_t *head)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

__global__ void ib_run_start_kernel(const uint32_t *head, const uint32_t *run_id_incl, uint32_t n, uint32_t *run_start)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && head[i]) run_start[run_id_incl[i] - 1] = i;
}

__global__ void ib_run_need_kernel(const uint32_t *run_start, uint32_t n_runs, uint32_t *need, unsigned long long *total)
{
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t mine = 0;
    if (r < n_runs) {
        uint32_t c = run_start[r + 1] - run_start[r];
        mine = c >= 2 ? c + 1 : 0;
        need[r] = mine;
    }
    // 64-bit grand total (block-aggregated) so the caller can refuse inputs whose 32-bit offsets would wrap
    __shared__ unsigned long long blk;
    if (threadIdx.x == 0) blk = 0;
    __syncthreads();
    unsigned w = __reduce_add_sync(FULL_MASK, mine);
    if ((threadIdx.x & 31) == 0 && w) atomicAdd(&blk, (unsigned long long)w);
    __syncthreads();
    if (threadIdx.x == 0 && blk) atomicAdd(total, blk);
}

__global__ void ib_fill_overflow_kernel(const uint32_t *run_id_incl, const uint32_t *run_start, const uint32_t *ovf_off, const uint32_t *vals,
                                        uint32_t n, uint32_t *overflow)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t r = run_id_incl[i] - 1;
    uint32_t s = run_start[r], c = run_start[r + 1] - s;
    if (c < 2) return;
    uint32_t off = ovf_off[r];
    if (i == s) overflow[off] = c;
    overflow[off + 1 + (i - s)] = vals[i];
}

// leader run of each canonical seed: counts seeds per hash table.  Runs are sorted by key, so a block sees only one or
// two distinct tables: aggregate per warp before touching the global counters.
__global__ void ib_count_tables_kernel(const unsigned long long *keys, const uint32_t *run_start, uint32_t n_runs, unsigned long long *table_count)
{
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    bool leader = false;
    uint32_t table = 0xffffffffu;
    if (r < n_runs) {
        uint64_t canon = keys[run_start[r]] >> 1;
        leader = !(r > 0 && (keys[run_start[r - 1]] >> 1) == canon);
        table = (uint32_t)(canon >> 32);
    }
    if (!leader) table = 0xffffffffu;
    unsigned peers = __match_any_sync(FULL_MASK, table);
    if (leader && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&table_count[table], (unsigned long long)__popc(peers));
}

__global__ void ib_insert_kernel(const unsigned long long *keys, const uint32_t *run_start, const uint32_t *ovf_off, const uint32_t *vals,
                                 uint32_t n_runs, uint32_t n_bases, HtEntry *tables, const uint64_t *table_start, const uint64_t *table_size)
{
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    const uint64_t ck = keys[run_start[r]];
    const uint64_t canon = ck >> 1;
    if (r > 0 && (keys[run_start[r - 1]] >> 1) == canon) return;  // not the leader
    uint32_t v[2] = {0xfffffffeu, 0xfffffffeu};                   // "unused, the other complement must exist"
    for (uint32_t q = r; q < n_runs && q < r + 2; q++) {
        uint64_t k2 = keys[run_start[q]];
        if ((k2 >> 1) != canon) break;
        uint32_t s = run_start[q], c = run_start[q + 1] - s;
        v[k2 & 1] = c == 1 ? vals[s] : n_bases + ovf_off[q];
    }
    const uint32_t hi = (uint32_t)(canon >> 32), lo = (uint32_t)canon;
    HtEntry *t = tables + table_start[hi];
    const uint64_t size = table_size[hi];
    uint64_t idx = ht_hash(lo) % size;
    uint64_t n = 0;
    for (;;) {  // SNAPHashTable probe sequence: +1, +4, +9, +16, then linear (HashTable.h:74-105)
        if (atomicCAS(&t[idx].v1, INVALID_LOC, v[0]) == INVALID_LOC) {
            t[idx].key = lo;
            t[idx].v2 = v[1];
            return;
        }
        n++;
        idx = (n < 5) ? (idx + n * n) % size : (idx + 1) % size;
    }
}
