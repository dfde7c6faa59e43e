// bgzf.h -- BGZF blocks (SAM/BAM specification section 4.1) as plain functions: the per-block logic of SURVEY.md section 8 row f4b's
// second half, the compressed container GzipWriterFilter::compressChunk produces for BAM output (SNAPLib/GzipDataWriter.cpp:281-340:
// one gzip member per chunk with the 6-byte "BC" extra field that holds the member's size, CRC-32 and ISIZE behind the data).
// What is reproduced is the CONTAINER and the CONTENT (every reader inflates the same bytes), not zlib's bytes: the deflate stream
// inside a block is this library's own -- one dynamic-Huffman block coding the chunk's byte histogram, no string matching -- or a stored
// block when that would not be smaller.  The kernels of bgzf_kernels.cuh call these from device code; tests/hostsim compiles the same
// header with g++ so that the logic is checked against zlib's inflate on a box without a GPU (test infrastructure only).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define BGZF_HD __host__ __device__ __forceinline__
#else
#define BGZF_HD static inline
#endif

#define BGZF_MAX_CHUNK 65024u  // input bytes per block (32 x 2032): header + stored fallback stay below the format's 65536-byte block
#define BGZF_HEADER 18u        // gzip header with the BC field
#define BGZF_FOOTER 8u         // CRC-32, ISIZE
#define BGZF_SYMS 257          // literals 0..255 and end-of-block
#define BGZF_MAX_BITS 15
#define BGZF_TABLE_BITS (3 + 5 + 5 + 4 + 19 * 3 + (BGZF_SYMS + 1) * 4)  // BFINAL/BTYPE, HLIT, HDIST, HCLEN, 19 code-length code lengths, 258 lengths

// ---- CRC-32 (the gzip polynomial, reflected) ------------------------------------------------------------------------------------
BGZF_HD uint32_t bgzf_crc_table_entry(uint32_t i)
{
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
    return c;
}
// the register after `n` more bytes, starting from `crc` (no pre- or post-conditioning)
BGZF_HD uint32_t bgzf_crc_raw(const uint32_t *table, uint32_t crc, const uint8_t *p, uint32_t n)
{
    for (uint32_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return crc;
}
// shift[k][j]: the register that bit j becomes after 2^k zero bytes (17 x 32 words).  Appending n zero bytes is linear in the register:
// crc(A || B) with initial value r0 = zeros(crc(A, r0), |B|) ^ crc(B, 0), which is how 32 lanes combine the CRCs of their slices.
BGZF_HD void bgzf_crc_shift_build(const uint32_t *table, uint32_t *shift /* [17][32] */)
{
    for (int j = 0; j < 32; j++) { const uint32_t v = 1u << j; shift[j] = table[v & 0xff] ^ (v >> 8); }
    for (int k = 1; k < 17; k++)
        for (int j = 0; j < 32; j++) {
            uint32_t v = shift[(k - 1) * 32 + j], r = 0;
            for (int b = 0; v; b++, v >>= 1) if (v & 1) r ^= shift[(k - 1) * 32 + b];
            shift[k * 32 + j] = r;
        }
}
BGZF_HD uint32_t bgzf_crc_zeros(const uint32_t *shift, uint32_t crc, uint32_t n_bytes)
{
    for (int k = 0; n_bytes; k++, n_bytes >>= 1)
        if (n_bytes & 1) {
            uint32_t v = crc, r = 0;
            for (int b = 0; v; b++, v >>= 1) if (v & 1) r ^= shift[k * 32 + b];
            crc = r;
        }
    return crc;
}

// ---- code lengths: a Huffman tree of the histogram, limited to 15 bits ----------------------------------------------------------
// hist[257] (hist[256], the end-of-block symbol, is forced to at least 1; at least one literal must be present).  Scratch: order[257],
// weight[513], parent[513] (uint32 each).  Two sorted queues (leaves by weight, internal nodes in creation order) give the tree in
// O(n); when it comes out deeper than 15 the weights are halved (not below 1) and the tree is rebuilt, which flattens it.
BGZF_HD void bgzf_code_lengths(const uint32_t *hist, uint8_t *len, uint32_t *order, uint32_t *weight, uint32_t *parent)
{
    uint32_t shift_down = 0;
    for (;;) {
        uint32_t n = 0;
        for (uint32_t s = 0; s < BGZF_SYMS; s++) {
            len[s] = 0;
            uint32_t w = s == 256 && hist[s] == 0 ? 1u : hist[s];
            if (!w) continue;
            w = (w >> shift_down) ? (w >> shift_down) : 1u;
            // insertion into the order sorted by (weight, symbol)
            uint32_t k = n++;
            while (k > 0 && weight[order[k - 1]] > w) { order[k] = order[k - 1]; k--; }
            order[k] = s;
            weight[s] = w;
        }
        if (n == 1) { len[order[0]] = 1; return; }  // (cannot happen with a literal and the end-of-block symbol; kept total)
        // leaves are nodes 0..256 (by symbol), internal nodes 257..; queue 1 walks `order`, queue 2 the internal nodes
        uint32_t q1 = 0, q2 = BGZF_SYMS, next = BGZF_SYMS;
        for (uint32_t made = 0; made + 1 < n; made++) {
            uint32_t pick[2];
            for (int t = 0; t < 2; t++) {
                const bool have1 = q1 < n, have2 = q2 < next;
                if (have1 && (!have2 || weight[order[q1]] <= weight[q2])) pick[t] = order[q1++];
                else pick[t] = q2++;
            }
            weight[next] = weight[pick[0]] + weight[pick[1]];
            parent[pick[0]] = parent[pick[1]] = next;
            next++;
        }
        const uint32_t root = next - 1;
        parent[root] = root;
        // depth of the internal nodes from the root down (a parent is always created after its children), then of the leaves
        uint32_t deepest = 0;
        weight[root] = 0;  // from here on weight[internal] holds the node's depth
        for (uint32_t v = root; v-- > BGZF_SYMS;) weight[v] = weight[parent[v]] + 1;
        for (uint32_t k = 0; k < n; k++) {
            const uint32_t s = order[k], d = weight[parent[s]] + 1;
            len[s] = (uint8_t)(d > 255 ? 255 : d);
            if (d > deepest) deepest = d;
        }
        if (deepest <= BGZF_MAX_BITS) return;
        shift_down++;
    }
}

// canonical codes (RFC 1951 section 3.2.2), bit-reversed so that they can be emitted least significant bit first
BGZF_HD void bgzf_assign_codes(const uint8_t *len, uint16_t *code)
{
    uint32_t count[BGZF_MAX_BITS + 2], next[BGZF_MAX_BITS + 2];
    for (int b = 0; b <= BGZF_MAX_BITS + 1; b++) count[b] = 0;
    for (int s = 0; s < BGZF_SYMS; s++) count[len[s]]++;
    count[0] = 0;
    uint32_t c = 0;
    next[0] = 0;
    for (int b = 1; b <= BGZF_MAX_BITS; b++) { c = (c + count[b - 1]) << 1; next[b] = c; }
    for (int s = 0; s < BGZF_SYMS; s++) {
        const uint32_t l = len[s];
        if (!l) { code[s] = 0; continue; }
        uint32_t v = next[l]++, r = 0;
        for (uint32_t b = 0; b < l; b++) { r = (r << 1) | (v & 1); v >>= 1; }
        code[s] = (uint16_t)r;
    }
}

// ---- the bit stream ---------------------------------------------------------------------------------------------------------------
struct BgzfBits {  // little-endian bit writer over bytes the caller has zeroed
    uint8_t *p;
    uint64_t acc;
    uint32_t n;  // bits in acc
};
BGZF_HD void bgzf_put(BgzfBits &w, uint32_t value, uint32_t bits)
{
    w.acc |= (uint64_t)value << w.n;
    w.n += bits;
    while (w.n >= 8) { *w.p++ = (uint8_t)w.acc; w.acc >>= 8; w.n -= 8; }
}
BGZF_HD void bgzf_flush(BgzfBits &w) { if (w.n) { *w.p++ = (uint8_t)w.acc; w.acc = 0; w.n = 0; } }

// The dynamic block's tables: BFINAL = 1, BTYPE = 2, 257 literal/length codes, one distance code (of length 0 -- RFC 1951: "one
// distance code of zero bits means that there are no distance codes used at all"), and the code-length alphabet used without its
// repeat symbols: symbols 0..15 get 4 bits each (a complete code), 16..18 none, so every one of the 258 lengths costs 4 bits.
BGZF_HD void bgzf_put_tables(BgzfBits &w, const uint8_t *len)
{
    bgzf_put(w, 1, 1); bgzf_put(w, 2, 2);
    bgzf_put(w, 0, 5);   // HLIT: 257 codes
    bgzf_put(w, 0, 5);   // HDIST: 1 code
    bgzf_put(w, 15, 4);  // HCLEN: all 19 code-length code lengths follow, in the order 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15
    for (int i = 0; i < 19; i++) bgzf_put(w, i < 3 ? 0 : 4, 3);
    // with sixteen 4-bit codes the canonical code of symbol v is v itself; reversed for the LSB-first stream
    for (int s = 0; s <= BGZF_SYMS; s++) {
        const uint32_t v = s < BGZF_SYMS ? len[s] : 0, r = ((v & 1) << 3) | ((v & 2) << 1) | ((v & 4) >> 1) | ((v & 8) >> 3);
        bgzf_put(w, r, 4);
    }
}

// bits of the dynamic block for this histogram (tables + symbols + end of block)
BGZF_HD uint64_t bgzf_dynamic_bits(const uint32_t *hist, const uint8_t *len)
{
    uint64_t bits = BGZF_TABLE_BITS + len[256];
    for (int s = 0; s < 256; s++) bits += (uint64_t)hist[s] * len[s];
    return bits;
}

BGZF_HD void bgzf_put_header(uint8_t *p, uint32_t block_size)
{  // SAM/BAM specification 4.1: ID1 ID2 CM FLG MTIME(4) XFL OS XLEN(2) SI1 SI2 SLEN(2) BSIZE(2); GzipDataWriter.cpp:296-312 sets time 0, xflags 0, os 0
    const uint8_t h[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 0, 6, 0, 'B', 'C', 2, 0};
    for (int i = 0; i < 16; i++) p[i] = h[i];
    p[16] = (uint8_t)(block_size - 1);
    p[17] = (uint8_t)((block_size - 1) >> 8);
}
BGZF_HD void bgzf_put_footer(uint8_t *p, uint32_t crc, uint32_t isize)
{
    for (int i = 0; i < 4; i++) { p[i] = (uint8_t)(crc >> (8 * i)); p[4 + i] = (uint8_t)(isize >> (8 * i)); }
}

// One whole block, serially (the specification the kernel is checked against): out must hold BGZF_HEADER + n + 5 + BGZF_FOOTER bytes
// and be zeroed.  Returns the block's size.  scratch: 257 + 257 + 513 + 513 uint32.
BGZF_HD uint32_t bgzf_block_serial(const uint32_t *crc_table, const uint8_t *in, uint32_t n, uint8_t *out, uint32_t *scratch)
{
    uint32_t *hist = scratch, *order = scratch + 257, *weight = order + 257, *parent = weight + 513;
    uint8_t len[BGZF_SYMS];
    uint16_t code[BGZF_SYMS];
    uint8_t *body = out + BGZF_HEADER;
    uint32_t body_bytes;
    bool stored = n == 0;
    if (!stored) {
        for (int s = 0; s < BGZF_SYMS; s++) hist[s] = 0;
        for (uint32_t i = 0; i < n; i++) hist[in[i]]++;
        hist[256] = 1;
        bgzf_code_lengths(hist, len, order, weight, parent);
        stored = (bgzf_dynamic_bits(hist, len) + 7) / 8 >= (uint64_t)n + 5;
    }
    if (stored) {  // BFINAL = 1, BTYPE = 0, padding to the byte, LEN, NLEN, the bytes
        body[0] = 1;
        body[1] = (uint8_t)n; body[2] = (uint8_t)(n >> 8); body[3] = (uint8_t)~n; body[4] = (uint8_t)(~n >> 8);
        for (uint32_t i = 0; i < n; i++) body[5 + i] = in[i];
        body_bytes = 5 + n;
    } else {
        bgzf_assign_codes(len, code);
        BgzfBits w = {body, 0, 0};
        bgzf_put_tables(w, len);
        for (uint32_t i = 0; i < n; i++) bgzf_put(w, code[in[i]], len[in[i]]);
        bgzf_put(w, code[256], len[256]);
        bgzf_flush(w);
        body_bytes = (uint32_t)(w.p - body);
    }
    const uint32_t total = BGZF_HEADER + body_bytes + BGZF_FOOTER;
    bgzf_put_header(out, total);
    bgzf_put_footer(body + body_bytes, bgzf_crc_raw(crc_table, 0xffffffffu, in, n) ^ 0xffffffffu, n);
    return total;
}
