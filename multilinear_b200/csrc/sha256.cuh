// SHA-256 (FIPS 180-4) for the Merkle tree and transcript of fr34za/multilinear
// (reference: `sha2` 0.10.8 call sites src/merkle_tree/mod.rs:110-116, 178-189; src/transcript.rs:23-38).
//
// Device side: fully unrolled 64-round compression with a 16-word rolling schedule, all in registers.
// Digests live in registers as the eight big-endian state words; memory holds the byte string, so
// loads/stores byte-swap (PRMT).  Specialisations used by the tree kernels:
//   sha256_leaf32   : 32-byte message (one ReedSolomonPair, src/fri/mod.rs:30-43) -> 1 compression
//   sha256_node64   : 64-byte message (left || right digest) -> 2 compressions, the 2nd over the
//                     constant padding block whose schedule folds to immediates
// Host side: a small streaming implementation for the Fiat-Shamir transcript (<= a few hundred bytes per call).
#pragma once
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

namespace mlb {

#define MLB_SHA_K_LIST                                                                                              \
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,         \
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,         \
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,         \
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,         \
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,         \
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,         \
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,         \
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u

__device__ static constexpr uint32_t kShaK[64] = {MLB_SHA_K_LIST};
static const uint32_t kShaKHost[64] = {MLB_SHA_K_LIST};

#define MLB_SHA_IV0 0x6a09e667u
#define MLB_SHA_IV1 0xbb67ae85u
#define MLB_SHA_IV2 0x3c6ef372u
#define MLB_SHA_IV3 0xa54ff53au
#define MLB_SHA_IV4 0x510e527fu
#define MLB_SHA_IV5 0x9b05688cu
#define MLB_SHA_IV6 0x1f83d9abu
#define MLB_SHA_IV7 0x5be0cd19u

__device__ __forceinline__ uint32_t sha_rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
__device__ __forceinline__ uint32_t sha_bswap(uint32_t x) { return __byte_perm(x, 0u, 0x0123); }
__device__ __forceinline__ uint32_t sha_S0(uint32_t x) { return sha_rotr(x, 2) ^ sha_rotr(x, 13) ^ sha_rotr(x, 22); }
__device__ __forceinline__ uint32_t sha_S1(uint32_t x) { return sha_rotr(x, 6) ^ sha_rotr(x, 11) ^ sha_rotr(x, 25); }
__device__ __forceinline__ uint32_t sha_s0(uint32_t x) { return sha_rotr(x, 7) ^ sha_rotr(x, 18) ^ (x >> 3); }
__device__ __forceinline__ uint32_t sha_s1(uint32_t x) { return sha_rotr(x, 17) ^ sha_rotr(x, 19) ^ (x >> 10); }
__device__ __forceinline__ uint32_t sha_ch(uint32_t e, uint32_t f, uint32_t g) { return (e & f) ^ (~e & g); }
__device__ __forceinline__ uint32_t sha_maj(uint32_t a, uint32_t b, uint32_t c) { return (a & b) ^ (a & c) ^ (b & c); }

// Rotations / shifts on the fma pipe: x * 2^(32-n) as a 64-bit product has x >> n in its high word and
// x << (32-n) in its low word, so  rotr(x, n) = IMAD(x, 2^(32-n), IMAD.HI(x, 2^(32-n)))  and  x >> n = IMAD.HI(x, 2^(32-n)).
// The multiplier comes from the constant bank so neither nvcc nor ptxas can turn it back into a funnel shift.
// This trades 1 alu-pipe instruction for 2 (rotate) or 1 (shift) fma-pipe instructions; the alu pipe is the bound.
__constant__ uint32_t kShaPow2[33] = {0u,          1u << 31, 1u << 30, 1u << 29, 1u << 28, 1u << 27, 1u << 26, 1u << 25, 1u << 24, 1u << 23, 1u << 22,
                                      1u << 21, 1u << 20, 1u << 19, 1u << 18, 1u << 17, 1u << 16, 1u << 15, 1u << 14, 1u << 13, 1u << 12,
                                      1u << 11, 1u << 10, 1u << 9,  1u << 8,  1u << 7,  1u << 6,  1u << 5,  1u << 4,  1u << 3,  1u << 2,
                                      1u << 1,  1u};  // kShaPow2[n] = 2^(32-n)
template <int N>
__device__ __forceinline__ uint32_t sha_shr_fma(uint32_t x) {
    uint32_t r;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(kShaPow2[N]));
    return r;
}
template <int N>
__device__ __forceinline__ uint32_t sha_rotr_fma(uint32_t x) {
    uint32_t r;
    asm("{\n\t.reg .u32 t;\n\tmul.hi.u32 t, %1, %2;\n\tmad.lo.u32 %0, %1, %2, t;\n\t}" : "=r"(r) : "r"(x), "r"(kShaPow2[N]));
    return r;
}
// NF = how many of the rotations go to the fma pipe (the rest are funnel shifts), SHRF = shift on the fma pipe
template <int NF> __device__ __forceinline__ uint32_t sha_S0_t(uint32_t x) {
    return (NF > 0 ? sha_rotr_fma<2>(x) : sha_rotr(x, 2)) ^ (NF > 1 ? sha_rotr_fma<13>(x) : sha_rotr(x, 13)) ^ (NF > 2 ? sha_rotr_fma<22>(x) : sha_rotr(x, 22));
}
template <int NF> __device__ __forceinline__ uint32_t sha_S1_t(uint32_t x) {
    return (NF > 0 ? sha_rotr_fma<6>(x) : sha_rotr(x, 6)) ^ (NF > 1 ? sha_rotr_fma<11>(x) : sha_rotr(x, 11)) ^ (NF > 2 ? sha_rotr_fma<25>(x) : sha_rotr(x, 25));
}
template <int NF, int SHRF> __device__ __forceinline__ uint32_t sha_s0_t(uint32_t x) {
    return (NF > 0 ? sha_rotr_fma<7>(x) : sha_rotr(x, 7)) ^ (NF > 1 ? sha_rotr_fma<18>(x) : sha_rotr(x, 18)) ^ (SHRF ? sha_shr_fma<3>(x) : (x >> 3));
}
template <int NF, int SHRF> __device__ __forceinline__ uint32_t sha_s1_t(uint32_t x) {
    return (NF > 0 ? sha_rotr_fma<17>(x) : sha_rotr(x, 17)) ^ (NF > 1 ? sha_rotr_fma<19>(x) : sha_rotr(x, 19)) ^ (SHRF ? sha_shr_fma<10>(x) : (x >> 10));
}

__device__ __forceinline__ void sha_iv(uint32_t st[8]) {
    st[0] = MLB_SHA_IV0; st[1] = MLB_SHA_IV1; st[2] = MLB_SHA_IV2; st[3] = MLB_SHA_IV3;
    st[4] = MLB_SHA_IV4; st[5] = MLB_SHA_IV5; st[6] = MLB_SHA_IV6; st[7] = MLB_SHA_IV7;
}

// Additions can be routed to the fma pipe.  SHA-256 is bound by the alu pipe (rotates = SHF, boolean functions =
// LOP3, about 1.0k such instructions per compression, 2 cycles each per SM sub-partition).  `x * 1 + y` with the 1
// read from the constant bank cannot be folded, so it is emitted as IMAD on the fma pipe.
// MASK bit set -> that group of additions is emitted as IMAD; clear -> left to ptxas (IADD3 / IMAD.IADD mix):
//   1 message schedule, 2 h + K + W, 4 t1 (+ch, +S1), 8 t2 = S0 + maj, 16 e = d + t1, 32 a = t1 + t2
// Measured on B200 (tools/shabench.py, profiles/r1_sha_add_routing.txt); the default is the fastest mix.
__constant__ uint32_t kShaOne = 1u;
template <bool FMA_ADD>
__device__ __forceinline__ uint32_t sha_add(uint32_t a, uint32_t b) {
    return FMA_ADD ? a * kShaOne + b : a + b;
}
#ifndef MLB_SHA_ADD_MASK
#define MLB_SHA_ADD_MASK 2
#endif

// One compression; w[16] holds the block as big-endian words and is clobbered (rolling schedule).
// ROT = 100 * (big-sigma rotations on the fma pipe, 0..3) + 10 * (small-sigma rotations, 0..2) + (small-sigma shift on fma, 0/1)
template <int MASK, int ROT = 0>
__device__ __forceinline__ void sha_compress_t(uint32_t st[8], uint32_t w[16]) {
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        if (i >= 16)
            w[i & 15] = sha_add<(MASK & 1) != 0>(sha_add<(MASK & 1) != 0>(w[i & 15], sha_s0_t<(ROT / 10) % 10, ROT % 10>(w[(i + 1) & 15])),
                                                 sha_add<(MASK & 1) != 0>(w[(i + 9) & 15], sha_s1_t<(ROT / 10) % 10, ROT % 10>(w[(i + 14) & 15])));
        // h + K + W first: it does not depend on this round's e, so it stays off the critical path
        uint32_t hkw = sha_add<(MASK & 2) != 0>(sha_add<(MASK & 2) != 0>(w[i & 15], kShaK[i]), h);
        uint32_t t1 = sha_add<(MASK & 4) != 0>(sha_add<(MASK & 4) != 0>(hkw, sha_ch(e, f, g)), sha_S1_t<ROT / 100>(e));
        uint32_t t2 = sha_add<(MASK & 8) != 0>(sha_S0_t<ROT / 100>(a), sha_maj(a, b, c));
        h = g; g = f; f = e; e = sha_add<(MASK & 16) != 0>(d, t1); d = c; c = b; b = a; a = sha_add<(MASK & 32) != 0>(t1, t2);
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}
#ifndef MLB_SHA_ROT
#define MLB_SHA_ROT 1
#endif
__device__ __forceinline__ void sha_compress(uint32_t st[8], uint32_t w[16]) { sha_compress_t<MLB_SHA_ADD_MASK, MLB_SHA_ROT>(st, w); }

// Compression over a constant block (the padding block that ends every 64-byte message): the whole message
// schedule is known at compile time, so round i only needs the immediate K[i] + W[i].
struct ShaKW {
    uint32_t v[64];
};
constexpr uint32_t sha_rotr_c(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
constexpr ShaKW sha_make_pad_kw(uint32_t message_bits) {
    constexpr uint32_t K[64] = {MLB_SHA_K_LIST};
    uint32_t w[64] = {};
    w[0] = 0x80000000u;
    w[15] = message_bits;
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = sha_rotr_c(w[i - 15], 7) ^ sha_rotr_c(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = sha_rotr_c(w[i - 2], 17) ^ sha_rotr_c(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    ShaKW r = {};
    for (int i = 0; i < 64; i++) r.v[i] = K[i] + w[i];
    return r;
}
__device__ static constexpr ShaKW kShaPad512 = sha_make_pad_kw(512u);

template <int MASK, int ROT = 0>
__device__ __forceinline__ void sha_compress_pad512_t(uint32_t st[8]) {
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        uint32_t hkw = sha_add<(MASK & 2) != 0>(h, kShaPad512.v[i]);
        uint32_t t1 = sha_add<(MASK & 4) != 0>(sha_add<(MASK & 4) != 0>(hkw, sha_ch(e, f, g)), sha_S1_t<ROT / 100>(e));
        uint32_t t2 = sha_add<(MASK & 8) != 0>(sha_S0_t<ROT / 100>(a), sha_maj(a, b, c));
        h = g; g = f; f = e; e = sha_add<(MASK & 16) != 0>(d, t1); d = c; c = b; b = a; a = sha_add<(MASK & 32) != 0>(t1, t2);
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}
// Code size matters for the tree kernels (one loop body holds one copy of each compression): a second, leaf-specialised copy
// (65 KB body) ran 14 % slower; rolling this function into 4 x 16 rounds with constants loaded at run time (31 KB body, under the
// 32 KB instruction-cache level) changed nothing measurable (5.90 vs 5.91 ms per commit) — 41 KB is not yet the limiter.
__device__ __forceinline__ void sha_compress_pad512(uint32_t st[8]) { sha_compress_pad512_t<MLB_SHA_ADD_MASK, MLB_SHA_ROT>(st); }

// SHA-256 of a 32-byte message given as 8 big-endian words.
__device__ __forceinline__ void sha256_leaf32(const uint32_t m[8], uint32_t out[8]) {
    uint32_t w[16] = {m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], 0x80000000u, 0u, 0u, 0u, 0u, 0u, 0u, 256u};
    sha_iv(out);
    sha_compress(out, w);
}

// SHA-256(left || right) for two 32-byte digests given as state words (hash_node, src/merkle_tree/mod.rs:184-189).
__device__ __forceinline__ void sha256_node64(const uint32_t l[8], const uint32_t r[8], uint32_t out[8]) {
    uint32_t w[16] = {l[0], l[1], l[2], l[3], l[4], l[5], l[6], l[7], r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]};
    sha_iv(out);
    sha_compress(out, w);
    sha_compress_pad512(out);  // constant padding block: schedule folded into immediates
}

// Field element (LE limbs as stored) -> the 4 big-endian message words of its 16 bytes.
__device__ __forceinline__ void sha_words_from_le(uint4 x, uint32_t* w) {
    w[0] = sha_bswap(x.x); w[1] = sha_bswap(x.y); w[2] = sha_bswap(x.z); w[3] = sha_bswap(x.w);
}
__device__ __forceinline__ void sha_store_digest(uint8_t* dst, const uint32_t st[8]) {
    uint4* p = reinterpret_cast<uint4*>(dst);
    p[0] = make_uint4(sha_bswap(st[0]), sha_bswap(st[1]), sha_bswap(st[2]), sha_bswap(st[3]));
    p[1] = make_uint4(sha_bswap(st[4]), sha_bswap(st[5]), sha_bswap(st[6]), sha_bswap(st[7]));
}
__device__ __forceinline__ void sha_load_digest(const uint8_t* src, uint32_t st[8]) {
    const uint4* p = reinterpret_cast<const uint4*>(src);
    uint4 a = p[0], b = p[1];
    st[0] = sha_bswap(a.x); st[1] = sha_bswap(a.y); st[2] = sha_bswap(a.z); st[3] = sha_bswap(a.w);
    st[4] = sha_bswap(b.x); st[5] = sha_bswap(b.y); st[6] = sha_bswap(b.z); st[7] = sha_bswap(b.w);
}

// Generic message of `len` bytes at an arbitrary (unaligned) address — used for `Merkle<T>::commit`
// over caller-defined items (src/merkle_tree/mod.rs:65-85) and batched leaves of odd item sizes.
// `next(i)` returns byte i of the message.
template <typename ByteFn>
__device__ inline void sha256_bytes(ByteFn next, size_t len, uint32_t out[8]) {
    sha_iv(out);
    size_t total_blocks = (len + 9 + 63) / 64;
    for (size_t blk = 0; blk < total_blocks; blk++) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
            uint32_t word = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                size_t pos = blk * 64 + (size_t)(4 * i + j);
                uint32_t byte = pos < len ? (uint32_t)next(pos) : (pos == len ? 0x80u : 0u);
                word = (word << 8) | byte;
            }
            w[i] = word;
        }
        if (blk == total_blocks - 1) {
            uint64_t bits = (uint64_t)len * 8ull;
            w[14] = (uint32_t)(bits >> 32);
            w[15] = (uint32_t)bits;
        }
        sha_compress(out, w);
    }
}

// ----------------------------------------------------------------------------- host (transcript only)
struct HostSha256 {
    uint32_t h[8];
    uint8_t buf[64];
    uint64_t len;
    HostSha256() { reset(); }
    void reset() {
        const uint32_t iv[8] = {MLB_SHA_IV0, MLB_SHA_IV1, MLB_SHA_IV2, MLB_SHA_IV3, MLB_SHA_IV4, MLB_SHA_IV5, MLB_SHA_IV6, MLB_SHA_IV7};
        memcpy(h, iv, sizeof iv);
        len = 0;
    }
    static uint32_t rr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void block(const uint8_t* p) {
        uint32_t w[64];
        for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
        for (int i = 16; i < 64; i++)
            w[i] = w[i - 16] + (rr(w[i - 15], 7) ^ rr(w[i - 15], 18) ^ (w[i - 15] >> 3)) + w[i - 7] + (rr(w[i - 2], 17) ^ rr(w[i - 2], 19) ^ (w[i - 2] >> 10));
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; i++) {
            uint32_t t1 = hh + (rr(e, 6) ^ rr(e, 11) ^ rr(e, 25)) + ((e & f) ^ (~e & g)) + kShaKHost[i] + w[i];
            uint32_t t2 = (rr(a, 2) ^ rr(a, 13) ^ rr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    void update(const void* data, size_t n) {
        const uint8_t* p = (const uint8_t*)data;
        size_t fill = (size_t)(len & 63);
        len += n;
        if (fill) {
            size_t take = 64 - fill < n ? 64 - fill : n;
            memcpy(buf + fill, p, take);
            p += take; n -= take; fill += take;
            if (fill < 64) return;
            block(buf);
        }
        while (n >= 64) { block(p); p += 64; n -= 64; }
        if (n) memcpy(buf, p, n);
    }
    // digest of everything absorbed so far; the running state is left untouched (Transcript::random clones, src/transcript.rs:23-29)
    void digest(uint8_t out[32]) const {
        HostSha256 c = *this;
        uint8_t pad[72];
        size_t fill = (size_t)(c.len & 63), padlen = (fill < 56 ? 56 : 120) - fill;
        uint64_t bits = c.len * 8;
        memset(pad, 0, sizeof pad);
        pad[0] = 0x80;
        c.update(pad, padlen);
        uint8_t lb[8];
        for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
        c.update(lb, 8);
        for (int i = 0; i < 8; i++) { out[4 * i] = c.h[i] >> 24; out[4 * i + 1] = c.h[i] >> 16; out[4 * i + 2] = c.h[i] >> 8; out[4 * i + 3] = c.h[i]; }
    }
};

}  // namespace mlb
