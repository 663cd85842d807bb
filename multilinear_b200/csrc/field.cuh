// Device arithmetic in the 128-bit prime field of fr34za/multilinear:
//   M = 2^128 - 45*2^40 + 1   (reference: src/ntt/mod.rs:35, Field128 = winter-math f128, src/field.rs:31)
// Elements are canonical (0 <= x < M), four 32-bit limbs, little-endian in memory exactly as the
// reference's `AsRef<[u8]>` exposes them (src/field.rs:33-38), so they can be hashed as stored.
//
// sm_100a integer pipes: 32x32->64 multiply-adds are IMAD.WIDE.U32 on the fma pipe with the carry
// chained through predicates (ptxas fuses the mad.lo.cc / madc.hi.cc pairs below), additions are
// IADD3.X on the alu pipe.  The 256-bit product is folded with 2^128 == c (mod M), c = 45*2^40 - 1,
// which keeps values in plain (non-Montgomery) form.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mlb {

struct fe {
    uint32_t v[4];
};

#define MLB_C0 0xFFFFFFFFu  // c = 2^128 - M = 0x2CFF_FFFFFFFF
#define MLB_C1 0x00002CFFu
#define MLB_M0 0x00000001u  // M limbs
#define MLB_M1 0xFFFFD300u
#define MLB_M2 0xFFFFFFFFu
#define MLB_M3 0xFFFFFFFFu

__device__ __forceinline__ fe fe_zero() { return fe{{0u, 0u, 0u, 0u}}; }
__device__ __forceinline__ fe fe_one() { return fe{{1u, 0u, 0u, 0u}}; }
__device__ __forceinline__ fe fe_from_u4(uint4 x) { return fe{{x.x, x.y, x.z, x.w}}; }
__device__ __forceinline__ uint4 fe_to_u4(fe a) { return make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]); }
__device__ __forceinline__ fe fe_load(const fe* p) { return fe_from_u4(*reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ fe fe_load_nc(const fe* p) { return fe_from_u4(__ldg(reinterpret_cast<const uint4*>(p))); }
__device__ __forceinline__ void fe_store(fe* p, fe a) { *reinterpret_cast<uint4*>(p) = fe_to_u4(a); }
__device__ __forceinline__ bool fe_eq(fe a, fe b) {
    return ((a.v[0] ^ b.v[0]) | (a.v[1] ^ b.v[1]) | (a.v[2] ^ b.v[2]) | (a.v[3] ^ b.v[3])) == 0u;
}
__device__ __forceinline__ bool fe_is_zero(fe a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0u; }

// BaseElement::new — one conditional subtraction of M (x - M == x + c mod 2^128)
__device__ __forceinline__ fe fe_new(fe x) {
    uint32_t z0, z1, z2, z3, g;
    asm("add.cc.u32 %0, %5, %9; addc.cc.u32 %1, %6, %10; addc.cc.u32 %2, %7, 0; addc.cc.u32 %3, %8, 0; addc.u32 %4, 0, 0;"
        : "=&r"(z0), "=&r"(z1), "=&r"(z2), "=&r"(z3), "=&r"(g)
        : "r"(x.v[0]), "r"(x.v[1]), "r"(x.v[2]), "r"(x.v[3]), "r"(MLB_C0), "r"(MLB_C1));
    return fe{{g ? z0 : x.v[0], g ? z1 : x.v[1], g ? z2 : x.v[2], g ? z3 : x.v[3]}};
}

__device__ __forceinline__ fe fe_add(fe a, fe b) {
    uint32_t s0, s1, s2, s3, k, z0, z1, z2, z3, g;
    asm("add.cc.u32 %0, %5, %9; addc.cc.u32 %1, %6, %10; addc.cc.u32 %2, %7, %11; addc.cc.u32 %3, %8, %12; addc.u32 %4, 0, 0;"
        : "=&r"(s0), "=&r"(s1), "=&r"(s2), "=&r"(s3), "=&r"(k)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]));
    asm("add.cc.u32 %0, %5, %9; addc.cc.u32 %1, %6, %10; addc.cc.u32 %2, %7, 0; addc.cc.u32 %3, %8, 0; addc.u32 %4, %11, 0;"
        : "=&r"(z0), "=&r"(z1), "=&r"(z2), "=&r"(z3), "=&r"(g)
        : "r"(s0), "r"(s1), "r"(s2), "r"(s3), "r"(MLB_C0), "r"(MLB_C1), "r"(k));
    // g != 0  <=>  a + b >= M
    return fe{{g ? z0 : s0, g ? z1 : s1, g ? z2 : s2, g ? z3 : s3}};
}

__device__ __forceinline__ fe fe_sub(fe a, fe b) {
    uint32_t d0, d1, d2, d3, br;
    asm("sub.cc.u32 %0, %5, %9; subc.cc.u32 %1, %6, %10; subc.cc.u32 %2, %7, %11; subc.cc.u32 %3, %8, %12; subc.u32 %4, 0, 0;"
        : "=&r"(d0), "=&r"(d1), "=&r"(d2), "=&r"(d3), "=&r"(br)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]));
    // br = 0xFFFFFFFF on borrow: add M back == subtract c (mod 2^128)
    asm("sub.cc.u32 %0, %0, %4; subc.cc.u32 %1, %1, %5; subc.cc.u32 %2, %2, 0; subc.u32 %3, %3, 0;"
        : "+r"(d0), "+r"(d1), "+r"(d2), "+r"(d3)
        : "r"(br & MLB_C0), "r"(br & MLB_C1));
    return fe{{d0, d1, d2, d3}};
}

__device__ __forceinline__ fe fe_neg(fe a) { return fe_sub(fe_zero(), a); }

// x / 2 : (x + (x odd ? M : 0)) >> 1
__device__ __forceinline__ fe fe_half(fe x) {
    uint32_t m = 0u - (x.v[0] & 1u), s0, s1, s2, s3, k;
    asm("add.cc.u32 %0, %5, %9; addc.cc.u32 %1, %6, %10; addc.cc.u32 %2, %7, %11; addc.cc.u32 %3, %8, %11; addc.u32 %4, 0, 0;"
        : "=&r"(s0), "=&r"(s1), "=&r"(s2), "=&r"(s3), "=&r"(k)
        : "r"(x.v[0]), "r"(x.v[1]), "r"(x.v[2]), "r"(x.v[3]), "r"(m & MLB_M0), "r"(m & MLB_M1), "r"(m));
    return fe{{__funnelshift_r(s0, s1, 1), __funnelshift_r(s1, s2, 1), __funnelshift_r(s2, s3, 1), __funnelshift_r(s3, k, 1)}};
}

// 4x4 limbs -> 8 limbs.  Even/odd column accumulators so every 32x32->64 product lands on a
// 64-bit aligned slot and the carries ride the mad chains (16 IMAD.WIDE.U32 + a 7-limb merge).
__device__ __forceinline__ void fe_mul_wide(uint32_t r[8], const fe& a, const fe& b) {
    uint32_t e0, e1, e2, e3, e4, e5, e6, e7, o0, o1, o2, o3, o4, o5, o6;
    asm("mul.lo.u32 %0, %4, %6; mul.hi.u32 %1, %4, %6; mul.lo.u32 %2, %5, %6; mul.hi.u32 %3, %5, %6;"
        : "=&r"(e0), "=&r"(e1), "=&r"(e2), "=&r"(e3) : "r"(a.v[0]), "r"(a.v[2]), "r"(b.v[0]));
    asm("mul.lo.u32 %0, %4, %6; mul.hi.u32 %1, %4, %6; mul.lo.u32 %2, %5, %6; mul.hi.u32 %3, %5, %6;"
        : "=&r"(o0), "=&r"(o1), "=&r"(o2), "=&r"(o3) : "r"(a.v[1]), "r"(a.v[3]), "r"(b.v[0]));
    asm("mad.lo.cc.u32 %0, %5, %7, %0; madc.hi.cc.u32 %1, %5, %7, %1; madc.lo.cc.u32 %2, %6, %7, %2; madc.hi.cc.u32 %3, %6, %7, %3; addc.u32 %4, 0, 0;"
        : "+r"(o0), "+r"(o1), "+r"(o2), "+r"(o3), "=&r"(o4) : "r"(a.v[0]), "r"(a.v[2]), "r"(b.v[1]));
    asm("mad.lo.cc.u32 %0, %4, %6, %0; madc.hi.cc.u32 %1, %4, %6, %1; madc.lo.cc.u32 %2, %5, %6, 0; madc.hi.u32 %3, %5, %6, 0;"
        : "+r"(e2), "+r"(e3), "=&r"(e4), "=&r"(e5) : "r"(a.v[1]), "r"(a.v[3]), "r"(b.v[1]));
    asm("mad.lo.cc.u32 %0, %5, %7, %0; madc.hi.cc.u32 %1, %5, %7, %1; madc.lo.cc.u32 %2, %6, %7, %2; madc.hi.cc.u32 %3, %6, %7, %3; addc.u32 %4, 0, 0;"
        : "+r"(e2), "+r"(e3), "+r"(e4), "+r"(e5), "=&r"(e6) : "r"(a.v[0]), "r"(a.v[2]), "r"(b.v[2]));
    asm("mad.lo.cc.u32 %0, %4, %6, %0; madc.hi.cc.u32 %1, %4, %6, %1; madc.lo.cc.u32 %2, %5, %6, %2; madc.hi.u32 %3, %5, %6, 0;"
        : "+r"(o2), "+r"(o3), "+r"(o4), "=&r"(o5) : "r"(a.v[1]), "r"(a.v[3]), "r"(b.v[2]));
    asm("mad.lo.cc.u32 %0, %5, %7, %0; madc.hi.cc.u32 %1, %5, %7, %1; madc.lo.cc.u32 %2, %6, %7, %2; madc.hi.cc.u32 %3, %6, %7, %3; addc.u32 %4, 0, 0;"
        : "+r"(o2), "+r"(o3), "+r"(o4), "+r"(o5), "=&r"(o6) : "r"(a.v[0]), "r"(a.v[2]), "r"(b.v[3]));
    asm("mad.lo.cc.u32 %0, %4, %6, %0; madc.hi.cc.u32 %1, %4, %6, %1; madc.lo.cc.u32 %2, %5, %6, %2; madc.hi.u32 %3, %5, %6, 0;"
        : "+r"(e4), "+r"(e5), "+r"(e6), "=&r"(e7) : "r"(a.v[1]), "r"(a.v[3]), "r"(b.v[3]));
    r[0] = e0;
    asm("add.cc.u32 %0, %7, %14; addc.cc.u32 %1, %8, %15; addc.cc.u32 %2, %9, %16; addc.cc.u32 %3, %10, %17; addc.cc.u32 %4, %11, %18; addc.cc.u32 %5, %12, %19; addc.u32 %6, %13, %20;"
        : "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7])
        : "r"(e1), "r"(e2), "r"(e3), "r"(e4), "r"(e5), "r"(e6), "r"(e7), "r"(o0), "r"(o1), "r"(o2), "r"(o3), "r"(o4), "r"(o5), "r"(o6));
}

// 256-bit value p -> canonical element.  Two folds with 2^128 == c, then one conditional subtract.
// v1: multiplies by the two limbs of c (12 wide multiply-adds on the half-rate IMAD.WIDE/IMAD.HI path).
__device__ __forceinline__ fe fe_reduce_wide_v1(const uint32_t p[8]) {
    uint32_t e0 = p[0], e1 = p[1], e2 = p[2], e3 = p[3], e4, e5, o0, o1, o2, o3, o4;
    const uint32_t c0 = MLB_C0, c1 = MLB_C1;
    // x[0..5] = lo + hi*c
    asm("mad.lo.cc.u32 %0, %5, %7, %0; madc.hi.cc.u32 %1, %5, %7, %1; madc.lo.cc.u32 %2, %6, %7, %2; madc.hi.cc.u32 %3, %6, %7, %3; addc.u32 %4, 0, 0;"
        : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(e3), "=&r"(e4) : "r"(p[4]), "r"(p[6]), "r"(c0));
    asm("mul.lo.u32 %0, %4, %6; mul.hi.u32 %1, %4, %6; mul.lo.u32 %2, %5, %6; mul.hi.u32 %3, %5, %6;"
        : "=&r"(o0), "=&r"(o1), "=&r"(o2), "=&r"(o3) : "r"(p[5]), "r"(p[7]), "r"(c0));
    asm("mad.lo.cc.u32 %0, %5, %7, %0; madc.hi.cc.u32 %1, %5, %7, %1; madc.lo.cc.u32 %2, %6, %7, %2; madc.hi.cc.u32 %3, %6, %7, %3; addc.u32 %4, 0, 0;"
        : "+r"(o0), "+r"(o1), "+r"(o2), "+r"(o3), "=&r"(o4) : "r"(p[4]), "r"(p[6]), "r"(c1));
    asm("mad.lo.cc.u32 %0, %4, %6, %0; madc.hi.cc.u32 %1, %4, %6, %1; madc.lo.cc.u32 %2, %5, %6, %2; madc.hi.u32 %3, %5, %6, 0;"
        : "+r"(e2), "+r"(e3), "+r"(e4), "=&r"(e5) : "r"(p[5]), "r"(p[7]), "r"(c1));
    uint32_t x0 = e0, x1, x2, x3, x4, x5;
    asm("add.cc.u32 %0, %5, %10; addc.cc.u32 %1, %6, %11; addc.cc.u32 %2, %7, %12; addc.cc.u32 %3, %8, %13; addc.u32 %4, %9, %14;"
        : "=&r"(x1), "=&r"(x2), "=&r"(x3), "=&r"(x4), "=&r"(x5)
        : "r"(e1), "r"(e2), "r"(e3), "r"(e4), "r"(e5), "r"(o0), "r"(o1), "r"(o2), "r"(o3), "r"(o4));
    // y = x[0..3] + (x4 + x5*2^32)*c ; (x4,x5) < 2^47 so the product is < 2^93
    uint32_t y0 = x0, y1 = x1, y2 = x2, y3 = x3, k1, q0, q1, q2, k2;
    asm("mad.lo.cc.u32 %0, %5, %7, %0; madc.hi.cc.u32 %1, %5, %7, %1; madc.lo.cc.u32 %2, %6, %8, %2; madc.hi.cc.u32 %3, %6, %8, %3; addc.u32 %4, 0, 0;"
        : "+r"(y0), "+r"(y1), "+r"(y2), "+r"(y3), "=&r"(k1) : "r"(x4), "r"(x5), "r"(c0), "r"(c1));
    asm("mul.lo.u32 %0, %3, %5; mul.hi.u32 %1, %3, %5; mad.lo.cc.u32 %0, %4, %6, %0; madc.hi.cc.u32 %1, %4, %6, %1; addc.u32 %2, 0, 0;"
        : "=&r"(q0), "=&r"(q1), "=&r"(q2) : "r"(x4), "r"(x5), "r"(c1), "r"(c0));
    asm("add.cc.u32 %0, %0, %4; addc.cc.u32 %1, %1, %5; addc.cc.u32 %2, %2, %6; addc.u32 %3, 0, 0;"
        : "+r"(y1), "+r"(y2), "+r"(y3), "=&r"(k2) : "r"(q0), "r"(q1), "r"(q2));
    // at most one wrap of 2^128 in total (true value < 2^128 + 2^93); wrapped value is tiny so +c cannot wrap again
    uint32_t m = 0u - (k1 + k2);
    asm("add.cc.u32 %0, %0, %4; addc.cc.u32 %1, %1, %5; addc.cc.u32 %2, %2, 0; addc.u32 %3, %3, 0;"
        : "+r"(y0), "+r"(y1), "+r"(y2), "+r"(y3) : "r"(m & c0), "r"(m & c1));
    return fe_new(fe{{y0, y1, y2, y3}});
}


// v2: c = 45*2^40 - 1 = (45*2^8) * 2^32 - 1, so hi*c = ((hi * 11520) << 32) - hi: one 14-bit constant, a whole-limb
// shift and a subtraction.  5 wide multiplies instead of 12; the fma pipe (IMAD.WIDE at half rate) is what bounds
// fe_mul, the extra IADD3s ride the alu pipe (model + carry analysis: tools/limb_model.py reduce_wide_v2).
#define MLB_K45 11520u  // 45 << 8
__device__ __forceinline__ fe fe_reduce_wide_v2(const uint32_t p[8]) {
    // a[1..5] = lo[1..3] + ((hi * 11520) << 32), as two carry-chained rows of IMAD.WIDE (even limbs of hi, then odd limbs);
    // the high halves of the products are < 2^14, so the last madc of each row cannot carry out
    uint32_t a1 = p[1], a2 = p[2], a3 = p[3], a4, a5, x0, x1, x2, x3, x4, x5;
    const uint32_t k = MLB_K45;
    asm("mad.lo.cc.u32 %0, %4, %6, %0; madc.hi.cc.u32 %1, %4, %6, %1; madc.lo.cc.u32 %2, %5, %6, %2; madc.hi.u32 %3, %5, %6, 0;"
        : "+r"(a1), "+r"(a2), "+r"(a3), "=&r"(a4) : "r"(p[4]), "r"(p[6]), "r"(k));
    asm("mad.lo.cc.u32 %0, %4, %6, %0; madc.hi.cc.u32 %1, %4, %6, %1; madc.lo.cc.u32 %2, %5, %6, %2; madc.hi.u32 %3, %5, %6, 0;"
        : "+r"(a2), "+r"(a3), "+r"(a4), "=&r"(a5) : "r"(p[5]), "r"(p[7]), "r"(k));
    // x[0..5] = lo + (t << 32) - hi  (>= 0 and < 2^174 as an integer, so the final borrow is always clear)
    asm("sub.cc.u32 %0, %6, %12; subc.cc.u32 %1, %7, %13; subc.cc.u32 %2, %8, %14; subc.cc.u32 %3, %9, %15; subc.cc.u32 %4, %10, 0; subc.u32 %5, %11, 0;"
        : "=&r"(x0), "=&r"(x1), "=&r"(x2), "=&r"(x3), "=&r"(x4), "=&r"(x5)
        : "r"(p[0]), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]));
    // second fold: (x4 + x5*2^32) < 2^46, times 11520 < 2^60
    const unsigned long long U = (unsigned long long)x4 * MLB_K45 + ((unsigned long long)(x5 * MLB_K45) << 32);
    uint32_t b1, b2, b3, k1, y0, y1, y2, y3, k2;
    asm("add.cc.u32 %0, %4, %7; addc.cc.u32 %1, %5, %8; addc.cc.u32 %2, %6, 0; addc.u32 %3, 0, 0;"
        : "=&r"(b1), "=&r"(b2), "=&r"(b3), "=&r"(k1) : "r"(x1), "r"(x2), "r"(x3), "r"((uint32_t)U), "r"((uint32_t)(U >> 32)));
    asm("sub.cc.u32 %0, %5, %9; subc.cc.u32 %1, %6, %10; subc.cc.u32 %2, %7, 0; subc.cc.u32 %3, %8, 0; subc.u32 %4, 0, 0;"
        : "=&r"(y0), "=&r"(y1), "=&r"(y2), "=&r"(y3), "=&r"(k2) : "r"(x0), "r"(b1), "r"(b2), "r"(b3), "r"(x4), "r"(x5));
    // k2 = 0 or 0xFFFFFFFF (borrow); the true value v = y + (k1 - borrow) * 2^128 is in [0, 2^128 + 2^107), so it wraps
    // 2^128 at most once.  z = y + c serves both cases: wrapped (v == y + c mod M, y < 2^107 so z < M and no carry) and
    // not wrapped (carry out of y + c  <=>  y >= M, and then z = y - M): the result is z if wrapped or carry, else y.
    uint32_t z0, z1, z2, z3, g;
    asm("add.cc.u32 %0, %5, %9; addc.cc.u32 %1, %6, %10; addc.cc.u32 %2, %7, 0; addc.cc.u32 %3, %8, 0; addc.u32 %4, %11, %12;"
        : "=&r"(z0), "=&r"(z1), "=&r"(z2), "=&r"(z3), "=&r"(g)
        : "r"(y0), "r"(y1), "r"(y2), "r"(y3), "r"(MLB_C0), "r"(MLB_C1), "r"(k1), "r"(k2));
    // g = k1 + k2 + carry (mod 2^32): k1 + k2 is 0 (no wrap: 0+0 or 1+0xFFFFFFFF) or 1 (wrap), and a wrap excludes a carry
    return fe{{g ? z0 : y0, g ? z1 : y1, g ? z2 : y2, g ? z3 : y3}};
}
#ifndef MLB_REDUCE_V
#define MLB_REDUCE_V 2
#endif
template <int V>
__device__ __forceinline__ fe fe_reduce_wide_t(const uint32_t p[8]) { return V == 1 ? fe_reduce_wide_v1(p) : fe_reduce_wide_v2(p); }
__device__ __forceinline__ fe fe_reduce_wide(const uint32_t p[8]) { return fe_reduce_wide_t<MLB_REDUCE_V>(p); }
template <int V>
__device__ __forceinline__ fe fe_mul_t(const fe& a, const fe& b) {
    uint32_t p[8];
    fe_mul_wide(p, a, b);
    return fe_reduce_wide_t<V>(p);
}

__device__ __forceinline__ fe fe_mul(const fe& a, const fe& b) {
    uint32_t p[8];
    fe_mul_wide(p, a, b);
    return fe_reduce_wide(p);
}
__device__ __forceinline__ fe fe_sqr(const fe& a) { return fe_mul(a, a); }

// a^e for a 64-bit exponent (table generation; exp as in FieldElement::exp, src/ntt/mod.rs:56-58)
__device__ inline fe fe_pow_u64(fe base, unsigned long long e) {
    fe r = fe_one();
    while (e) {
        if (e & 1ull) r = fe_mul(r, base);
        base = fe_sqr(base);
        e >>= 1;
    }
    return r;
}

// Lazy sum of products: 288-bit accumulator (9 limbs), reduced once at the end.
struct fe_acc {
    uint32_t w[9];
};
__device__ __forceinline__ void acc_zero(fe_acc& s) {
#pragma unroll
    for (int i = 0; i < 9; i++) s.w[i] = 0u;
}
__device__ __forceinline__ void acc_add_wide(fe_acc& s, const uint32_t p[8]) {
    asm("add.cc.u32 %0, %0, %9; addc.cc.u32 %1, %1, %10; addc.cc.u32 %2, %2, %11; addc.cc.u32 %3, %3, %12; "
        "addc.cc.u32 %4, %4, %13; addc.cc.u32 %5, %5, %14; addc.cc.u32 %6, %6, %15; addc.cc.u32 %7, %7, %16; addc.u32 %8, %8, 0;"
        : "+r"(s.w[0]), "+r"(s.w[1]), "+r"(s.w[2]), "+r"(s.w[3]), "+r"(s.w[4]), "+r"(s.w[5]), "+r"(s.w[6]), "+r"(s.w[7]), "+r"(s.w[8])
        : "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]));
}
__device__ __forceinline__ void acc_mul_add(fe_acc& s, const fe& a, const fe& b) {
    uint32_t p[8];
    fe_mul_wide(p, a, b);
    acc_add_wide(s, p);
}
// limbs 0..7 fold as a 256-bit value; the overflow limb w[8] counts 2^256 == c*c (mod M)
__device__ __forceinline__ fe acc_reduce(const fe_acc& s) {
    fe lo = fe_reduce_wide(s.w);
    fe c = fe{{MLB_C0, MLB_C1, 0u, 0u}};
    fe top = fe{{s.w[8], 0u, 0u, 0u}};
    fe t = fe_mul(fe_mul(top, c), c);
    return fe_add(lo, t);
}

}  // namespace mlb
