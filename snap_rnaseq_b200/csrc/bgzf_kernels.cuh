// bgzf_kernels.cuh -- BGZF blocks on the device (the compressed container; bgzf.h holds the per-block
// logic and says what is and is not reproduced).  One warp per block of at most BGZF_MAX_CHUNK input bytes:
//   histogram (shared-memory atomics, all lanes) -> code lengths and codes (leader lane: 257 symbols) -> every lane codes its own
//   contiguous slice of the chunk at the bit offset a warp scan of the slices' bit counts gives it (atomicOr into the zeroed slot, so
//   the words two slices share need no care) -> CRC-32 of the chunk from the 32 slice registers (bgzf_crc_zeros) -> header / footer.
// Blocks are built in fixed 64 KiB slots; bgzf_pack_kernel then moves them to their final offsets (a scan of the sizes in between).
#pragma once
#include "bgzf.h"

#define BGZF_SLOT 65536u
#define BGZF_WARPS 4

struct BgzfArgs {
    const uint8_t *in;
    unsigned long long n_bytes;
    uint32_t chunk, n_blocks;
    uint8_t *slots;       // [n_blocks][BGZF_SLOT], zeroed
    unsigned long long *sizes;  // [n_blocks + 1]
    const uint32_t *crc_table;  // [256]
    const uint32_t *crc_shift;  // [17][32]
    uint32_t *work;
};

struct BgzfSm {
    uint32_t hist[BGZF_SYMS + 3];
    uint32_t order[BGZF_SYMS + 3];
    uint32_t weight[516];
    uint32_t parent[516];
    uint16_t code[BGZF_SYMS + 3];
    uint8_t len[BGZF_SYMS + 3];
    uint32_t tables[(BGZF_TABLE_BITS + 31) / 32 + 1];  // the block's tables as the serial writer lays them out, OR-ed into the slot by the warp
    uint32_t stored;
};

// bits [pos, pos + nbits) of the slot, counted from its byte 16 (the aligned word that holds BSIZE and the first two body bytes)
__device__ __forceinline__ void bgzf_or_bits(uint32_t *words, unsigned long long pos, unsigned long long value, uint32_t nbits)
{
    if (!nbits) return;
    const uint32_t w = (uint32_t)(pos >> 5), sh = (uint32_t)(pos & 31);
    const unsigned long long lo = value << sh;
    atomicOr(&words[w], (uint32_t)lo);
    if (sh + nbits > 32) atomicOr(&words[w + 1], (uint32_t)(lo >> 32));  // nbits <= 32: never a third word
}

__global__ void __launch_bounds__(BGZF_WARPS * 32) bgzf_block_kernel(const BgzfArgs a)
{
    __shared__ BgzfSm sms[BGZF_WARPS];
    BgzfSm *sm = &sms[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    #pragma unroll 1
    for (;;) {
        uint32_t blk = 0;
        if (lane == 0) blk = atomicAdd(a.work, 1u);
        blk = __shfl_sync(0xffffffffu, blk, 0);
        if (blk >= a.n_blocks) break;
        const unsigned long long lo = (unsigned long long)blk * a.chunk;
        const uint32_t n = (uint32_t)min((unsigned long long)a.chunk, a.n_bytes - lo);
        const uint8_t *in = a.in + lo;
        uint8_t *slot = a.slots + (size_t)blk * BGZF_SLOT;
        uint32_t *words = (uint32_t *)(slot + 16);
        // this lane's slice
        const uint32_t per = (n + 31) / 32, s_lo = min(n, (uint32_t)lane * per), s_hi = min(n, s_lo + per);
        for (int s = lane; s < BGZF_SYMS; s += 32) sm->hist[s] = 0;
        __syncwarp();
        #pragma unroll 1
        for (uint32_t i = s_lo; i < s_hi; i++) atomicAdd(&sm->hist[in[i]], 1u);
        __syncwarp();
        if (lane == 0) {
            sm->hist[256] = 1;
            sm->stored = n == 0;
            if (n) bgzf_code_lengths(sm->hist, sm->len, sm->order, sm->weight, sm->parent);
        }
        __syncwarp();
        uint32_t body_bytes;
        if (n) {  // dynamic block or stored: whichever is smaller (all lanes evaluate the same sum)
            unsigned long long bits = 0;
            for (int s = lane; s < 256; s += 32) bits += (unsigned long long)sm->hist[s] * sm->len[s];
            #pragma unroll
            for (int o = 16; o; o >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, o);
            bits += BGZF_TABLE_BITS + sm->len[256];
            if (lane == 0) sm->stored = (bits + 7) / 8 >= (unsigned long long)n + 5;
        }
        __syncwarp();
        if (sm->stored) {
            if (lane == 0) { slot[18] = 1; slot[19] = (uint8_t)n; slot[20] = (uint8_t)(n >> 8); slot[21] = (uint8_t)~n; slot[22] = (uint8_t)(~n >> 8); }
            #pragma unroll 1
            for (uint32_t i = lane; i < n; i += 32) slot[23 + i] = in[i];
            body_bytes = 5 + n;
        } else {
            for (int k = lane; k < (int)(sizeof(sm->tables) / 4); k += 32) sm->tables[k] = 0;
            __syncwarp();
            if (lane == 0) {
                bgzf_assign_codes(sm->len, sm->code);
                BgzfBits w = {(uint8_t *)sm->tables, 0, 0};
                bgzf_put_tables(w, sm->len);
                bgzf_flush(w);
            }
            __syncwarp();
            // everything in the body goes in by atomicOr: the tables end in the middle of a byte that the first symbols share
            for (int k = lane; k < (int)((BGZF_TABLE_BITS + 31) / 32); k += 32) bgzf_or_bits(words, 16ull + 32ull * k, sm->tables[k], 32);
            uint32_t my_bits = 0;
            #pragma unroll 1
            for (uint32_t i = s_lo; i < s_hi; i++) my_bits += sm->len[in[i]];
            uint32_t incl = my_bits;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            const uint32_t total_bits = __shfl_sync(0xffffffffu, incl, 31);
            unsigned long long pos = 16ull + BGZF_TABLE_BITS + (incl - my_bits);
            unsigned long long acc = 0;
            uint32_t nacc = 0;
            #pragma unroll 1
            for (uint32_t i = s_lo; i < s_hi; i++) {
                const uint32_t b = in[i], l = sm->len[b];
                acc |= (unsigned long long)sm->code[b] << nacc;
                nacc += l;
                if (nacc >= 32) { bgzf_or_bits(words, pos, acc & 0xffffffffull, 32); pos += 32; acc >>= 32; nacc -= 32; }
            }
            bgzf_or_bits(words, pos, acc, nacc);
            if (lane == 31) bgzf_or_bits(words, 16ull + BGZF_TABLE_BITS + total_bits, sm->code[256], sm->len[256]);
            body_bytes = (uint32_t)((BGZF_TABLE_BITS + (unsigned long long)total_bits + sm->len[256] + 7) / 8);
        }
        // CRC-32 of the chunk from the slices' registers
        uint32_t r = lane == 0 ? 0xffffffffu : 0u;
        #pragma unroll 1
        for (uint32_t i = s_lo; i < s_hi; i++) r = __ldg(&a.crc_table[(r ^ in[i]) & 0xff]) ^ (r >> 8);
        r = bgzf_crc_zeros(a.crc_shift, r, n - s_hi);
        #pragma unroll
        for (int o = 16; o; o >>= 1) r ^= __shfl_xor_sync(0xffffffffu, r, o);
        __syncwarp();
        if (lane == 0) {
            const uint32_t total = BGZF_HEADER + body_bytes + BGZF_FOOTER;
            // bytes 16 and 17 share a word with body bits that were OR-ed in: OR the size in as well
            uint8_t h[18];
            bgzf_put_header(h, total);
            for (int i = 0; i < 16; i++) slot[i] = h[i];
            atomicOr(&words[0], (uint32_t)h[16] | ((uint32_t)h[17] << 8));
            a.sizes[blk] = total;
        }
        __syncwarp();
        if (lane == 0) bgzf_put_footer(slot + BGZF_HEADER + body_bytes, r ^ 0xffffffffu, n);
        __syncwarp();
    }
}

// slots -> the packed stream; one warp per block
__global__ void __launch_bounds__(256) bgzf_pack_kernel(const uint8_t *slots, const unsigned long long *offsets, uint32_t n_blocks, uint8_t *out)
{
    const uint32_t blk = (uint32_t)(((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (blk >= n_blocks) return;
    const unsigned long long o = offsets[blk];
    const uint32_t n = (uint32_t)(offsets[blk + 1] - o);
    const uint8_t *src = slots + (size_t)blk * BGZF_SLOT;
    for (uint32_t i = lane; i < n; i += 32) out[o + i] = src[i];
}
