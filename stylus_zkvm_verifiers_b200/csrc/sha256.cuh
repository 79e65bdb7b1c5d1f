// SHA-256 (FIPS 180-4) for the per-proof public-input hashing the reference does with the `sha2`
// crate: ReceiptClaim / Output digests (/root/reference/contracts/src/risc0/types.rs:62-95) and
// SP1 hash_public_values (/root/reference/contracts/src/sp1/types.rs:34-38).  One proof per thread.
#pragma once
#include <stdint.h>
#include "bn254.cuh"

namespace zkv {

#define ZKV_SHA_K_VALUES \
 \
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be, \
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, \
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, \
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, \
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, \
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2
// device code reads the __constant__ copy; host code (selector / vk digest, once per handle) its own copy
ZKV_CONST uint32_t SHA_K[64] = {ZKV_SHA_K_VALUES};
#if defined(__CUDACC__)
static const uint32_t SHA_K_HOST[64] = {ZKV_SHA_K_VALUES};
#endif
#if defined(__CUDACC__) && !defined(__CUDA_ARCH__)
#define ZKV_SHA_K_AT(i) SHA_K_HOST[i]
#else
#define ZKV_SHA_K_AT(i) SHA_K[i]
#endif

ZKV_HD ZKV_INLINE uint32_t ror32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

ZKV_HD ZKV_INLINE void sha256_init(uint32_t h[8]) {
    h[0] = 0x6a09e667; h[1] = 0xbb67ae85; h[2] = 0x3c6ef372; h[3] = 0xa54ff53a;
    h[4] = 0x510e527f; h[5] = 0x9b05688c; h[6] = 0x1f83d9ab; h[7] = 0x5be0cd19;
}
// one compression; w[16] = message block as big-endian words (clobbered)
ZKV_HD inline void sha256_compress(uint32_t h[8], uint32_t w[16]) {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            uint32_t s0 = ror32(w15, 7) ^ ror32(w15, 18) ^ (w15 >> 3), s1 = ror32(w2, 17) ^ ror32(w2, 19) ^ (w2 >> 10);
            w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
        }
        uint32_t t1 = hh + (ror32(e, 6) ^ ror32(e, 11) ^ ror32(e, 25)) + ((e & f) ^ (~e & g)) + ZKV_SHA_K_AT(i) + w[i & 15];
        uint32_t t2 = (ror32(a, 2) ^ ror32(a, 13) ^ ror32(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
ZKV_HD ZKV_INLINE uint32_t load_be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

// general message (used for SP1 public values and on the host for tags / selectors)
ZKV_HD inline void sha256_msg(uint32_t h[8], const uint8_t* msg, size_t len) {
    sha256_init(h);
    uint32_t w[16];
    size_t i = 0;
    for (; i + 64 <= len; i += 64) { for (int k = 0; k < 16; k++) w[k] = load_be32(msg + i + 4 * k); sha256_compress(h, w); }
    size_t rem = len - i;
    // tail: remaining bytes + 0x80 + zero pad + 64-bit length, in one or two blocks
    for (int blk = 0; blk < 2; blk++) {
        for (int k = 0; k < 16; k++) {
            uint32_t word = 0;
            for (int b = 0; b < 4; b++) {
                size_t pos = (size_t)blk * 64 + 4 * k + b;
                uint32_t byte = pos < rem ? msg[i + pos] : (pos == rem ? 0x80u : 0u);
                word = (word << 8) | byte;
            }
            w[k] = word;
        }
        bool last = (blk == 1) || (rem + 9 <= 64);
        if (last) { uint64_t bits = (uint64_t)len * 8; w[14] = (uint32_t)(bits >> 32); w[15] = (uint32_t)bits; }
        sha256_compress(h, w);
        if (last) break;
    }
}
ZKV_HD ZKV_INLINE void sha256_words_to_bytes(uint8_t out[32], const uint32_t h[8]) {
    for (int k = 0; k < 8; k++) { out[4 * k] = h[k] >> 24; out[4 * k + 1] = h[k] >> 16; out[4 * k + 2] = h[k] >> 8; out[4 * k + 3] = h[k]; }
}

// RISC Zero claim digest for ReceiptClaim::ok(image_id, journal_digest) (risc0/types.rs:44-95).
//   out   = SHA256(tagO || journal || 0^32 || 0x0200)                               (98 B, 2 blocks)
//   claim = SHA256(tagC || 0^32 || image_id || SYS0 || out || 0^4 || 0^4 || 0x0400) (170 B, 3 blocks)
// tag_out[8]    = SHA256("risc0.Output") as words; claim_mid[8] = state after block0 = tagC || 0^32
// (both per-handle constants computed on the host); sys0[8] = SYSTEM_STATE_ZERO_DIGEST words.
ZKV_HD inline void risc0_claim_digest(uint32_t claim[8], const uint32_t image_id[8], const uint32_t journal[8],
                                      const uint32_t tag_out[8], const uint32_t claim_mid[8], const uint32_t sys0[8]) {
    uint32_t h[8], w[16];
    sha256_init(h);
    for (int k = 0; k < 8; k++) { w[k] = tag_out[k]; w[8 + k] = journal[k]; }
    sha256_compress(h, w);
    for (int k = 0; k < 16; k++) w[k] = 0;
    w[8] = 0x02008000u; w[15] = 98 * 8;
    sha256_compress(h, w);
    for (int k = 0; k < 8; k++) claim[k] = claim_mid[k];
    for (int k = 0; k < 8; k++) { w[k] = image_id[k]; w[8 + k] = sys0[k]; }
    sha256_compress(claim, w);
    for (int k = 0; k < 8; k++) w[k] = h[k];
    w[8] = 0; w[9] = 0; w[10] = 0x04008000u; w[11] = 0; w[12] = 0; w[13] = 0; w[14] = 0; w[15] = 170 * 8;
    sha256_compress(claim, w);
}

}  // namespace zkv
