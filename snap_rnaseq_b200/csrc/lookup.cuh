// lookup.cuh -- seed packing and genome-index probes.
//
// Replaces Seed::Seed / DoesTextRepresentASeed (SNAPLib/Seed.h:38-51, Seed.cpp:29-42), SNAPHashTable::Lookup
// (SNAPLib/HashTable.h:74-105) and GenomeIndex::lookupSeed / fillInLookedUpResults
// (SNAPLib/GenomeIndex.cpp:971-1086, full-range form).
//
// One lane handles one seed: the lanes of a warp probe up to 32 seeds of a read at once, so the dependent
// chain "12-byte table entry -> overflow count word" of every seed is in flight together instead of one
// after the other as on the CPU (where lookupSeed is 41-45 % of the run time).
#pragma once
#include "common.cuh"

// A=0 G=1 C=2 T=3 (SNAPLib/Tables.cpp:36-42); anything else is not a seed base
__device__ __forceinline__ int base2(uint8_t c) { return c == 'A' ? 0 : c == 'G' ? 1 : c == 'C' ? 2 : c == 'T' ? 3 : -1; }

// rcTranslationTable (SNAPLib/BaseAligner.cpp:148-152); bytes the reference leaves unspecified map to 0
__device__ __forceinline__ uint8_t rc_base(uint8_t c)
{
    switch (c) {
        case 'A': return 'T';
        case 'C': return 'G';
        case 'G': return 'C';
        case 'T': return 'A';
        case 'N': return 'N';
    }
    return 0;
}

// first base in the most significant bits; reverse complement built alongside
__device__ __forceinline__ bool pack_seed(const uint8_t *text, uint32_t len, uint64_t *fwd, uint64_t *rc)
{
    uint64_t f = 0, r = 0;
    bool ok = true;
    #pragma unroll 1
    for (uint32_t i = 0; i < len; i++) {
        int v = base2(text[i]);
        ok &= v >= 0;
        f |= (uint64_t)(v & 3) << ((len - i - 1) * 2);
        r |= (uint64_t)((v & 3) ^ 3) << (i * 2);
    }
    *fwd = f;
    *rc = r;
    return ok;
}

__device__ __forceinline__ uint32_t ht_hash(uint32_t key)
{  // MurmurHash3 finalizer, HashTable.h:60-72
    key ^= key >> 16; key *= 0x85ebca6bu; key ^= key >> 13; key *= 0xc2b2ae35u; key ^= key >> 16;
    return key;
}

struct HitList {
    const uint32_t *hits;  // device pointer: into a table entry (singleton) or the overflow table
    uint32_t n;
};

__device__ __forceinline__ HitList resolve_hits(const DevIndex &ix, const uint32_t *sub)
{
    HitList r = {nullptr, 0};
    uint32_t v = __ldg(sub);
    if (v < ix.n_bases) { r.n = 1; r.hits = sub; }
    else if (v != 0xfffffffeu) {
        uint32_t off = v - ix.n_bases;
        r.n = __ldg(&ix.overflow[off]);
        r.hits = &ix.overflow[off + 1];
    }
    return r;
}

// GenomeIndex::lookupSeed.  probes (optional) counts table slots examined.
__device__ __forceinline__ void lookup_seed(const DevIndex &ix, uint64_t fwd, uint64_t rc, HitList out[2], uint32_t *probes)
{
    bool swapped = (int64_t)fwd > (int64_t)rc;
    uint64_t s = swapped ? rc : fwd;
    uint32_t hi = (uint32_t)(s >> 32), lo = (uint32_t)s;
    out[0].n = out[1].n = 0;
    out[0].hits = out[1].hits = nullptr;
    const HtEntry *t = ix.tables + ix.table_start[hi];
    const uint64_t size = ix.table_size[hi];
    uint32_t np = 1;
    const HtEntry *e;
    uint32_t key, v1;
    if (size <= 0xffffffffull) {  // every real table: 32-bit remainders instead of the 64-bit software routine
        const uint32_t sz = (uint32_t)size;
        uint32_t idx = ht_hash(lo) % sz;
        e = &t[idx];
        key = __ldg(&e->key); v1 = __ldg(&e->v1);
        if (!(key == lo && v1 != INVALID_LOC)) {
            uint32_t n = 0;
            #pragma unroll 1
            do {
                n++;
                if (n > sz && n - sz > 5) { e = nullptr; break; }
                const uint32_t step = n < 5 ? n * n : 1u;  // +1,+4,+9,+16, then linear (HashTable.h:74-105)
                idx = (uint32_t)(((uint64_t)idx + step) % sz);
                e = &t[idx];
                key = __ldg(&e->key);
                v1 = __ldg(&e->v1);
                np++;
            } while (key != lo && v1 != INVALID_LOC);
            if (e && v1 == INVALID_LOC) e = nullptr;
        }
    } else {
        uint64_t idx = ht_hash(lo) % size;
        e = &t[idx];
        key = __ldg(&e->key); v1 = __ldg(&e->v1);
        if (!(key == lo && v1 != INVALID_LOC)) {
            uint64_t n = 0;
            #pragma unroll 1
            do {
                n++;
                if (n > size + 5) { e = nullptr; break; }
                idx = (n < 5) ? (idx + n * n) % size : (idx + 1) % size;
                e = &t[idx];
                key = __ldg(&e->key);
                v1 = __ldg(&e->v1);
                np++;
            } while (key != lo && v1 != INVALID_LOC);
            if (e && v1 == INVALID_LOC) e = nullptr;
        }
    }
    if (probes) *probes = np;
    if (!e) return;
    out[0] = resolve_hits(ix, swapped ? &e->v2 : &e->v1);
    if (fwd == rc) out[1] = out[0];
    else out[1] = resolve_hits(ix, swapped ? &e->v1 : &e->v2);
}

// GetWrappedNextSeedToTest (SNAPLib/SeedSequencer.h:28-287) as data: row = seedLen-16.
__device__ const uint8_t WRAP_ORDER[10][25] = {
    {0, 8, 4, 12, 2, 6, 10, 14, 1, 3, 5, 7, 9, 11, 13, 15},
    {0, 8, 4, 12, 2, 6, 10, 14, 1, 3, 5, 7, 9, 11, 13, 15, 16},
    {0, 9, 4, 13, 2, 6, 11, 15, 1, 3, 5, 7, 8, 10, 12, 14, 16, 17},
    {0, 10, 4, 14, 2, 6, 8, 12, 16, 18, 1, 3, 5, 7, 9, 11, 13, 15, 17},
    {0, 10, 5, 15, 2, 7, 12, 17, 3, 9, 11, 13, 19, 1, 4, 6, 8, 14, 18, 16},
    {0, 11, 6, 16, 3, 9, 13, 17, 18, 2, 5, 8, 15, 20, 1, 4, 7, 10, 12, 14, 19},
    {0, 11, 6, 16, 3, 9, 14, 19, 2, 7, 12, 17, 20, 4, 1, 10, 13, 15, 18, 21, 5, 8},
    {0, 12, 6, 17, 3, 9, 20, 14, 1, 4, 7, 10, 15, 18, 21, 4, 2, 5, 11, 16, 19, 22, 8},  // sic: 4 appears twice
    {0, 12, 6, 18, 3, 15, 21, 9, 1, 13, 19, 7, 16, 4, 22, 10, 2, 14, 20, 5, 17, 8, 23, 11},
    {0, 13, 6, 19, 3, 16, 22, 9, 11, 1, 14, 7, 20, 4, 17, 23, 2, 15, 5, 21, 8, 24, 10, 18, 12},
};
__device__ __forceinline__ uint32_t wrapped_seed(uint32_t seed_len, uint32_t wrap) { return WRAP_ORDER[seed_len - 16][wrap]; }
