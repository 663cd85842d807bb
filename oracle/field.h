/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the
 * product; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.
 *
 * PARITY UNPINNED: the reference (Rust, /root/reference) cannot be built in
 * this image (no cargo/rustc, winter-math 0.12.0 and sha2 0.10.8 are not
 * vendored) and none of its 16 tests asserts a concrete value.  This file is a
 * CPU restatement; it is cross-checked against an independent Python big-int
 * restatement (oracle/pyref.py) and SURVEY.md §8c's provisional vectors.
 *
 * Field: winter-math 0.12.0 `math::fields::f128::BaseElement` as wrapped by
 * the reference's Field128 (src/field.rs:31).  Published semantics restated:
 *   - M = 2^128 - 45*2^40 + 1 (literal at src/ntt/mod.rs:35), values canonical
 *     in [0, M), stored as a little-endian u128 (src/field.rs:33-38).
 *   - new(x) does ONE conditional subtraction (enough since 2^128 < 2M).
 *   - inv(0) = 0; exp is square-and-multiply.
 */
#ifndef ORACLE_FIELD_H
#define ORACLE_FIELD_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef u128 fe;

#define FE_M ((((u128)0xFFFFFFFFFFFFFFFFULL) << 64) | (u128)0xFFFFD30000000001ULL)
/* c = 2^128 - M = 45*2^40 - 1, so 2^128 == c (mod M) */
#define FE_C ((uint64_t)0x2CFFFFFFFFFFULL)

static inline fe fe_new(u128 x) { return x >= FE_M ? x - FE_M : x; } /* BaseElement::new */
/* src/field.rs:144-154: From<i32>/From<i64> cast with sign extension, then new() */
static inline fe fe_from_i64(int64_t v) { return fe_new((u128)(__int128)v); }
static inline fe fe_load(const uint8_t *p) { fe x; memcpy(&x, p, 16); return x; }
static inline void fe_store(uint8_t *p, fe x) { memcpy(p, &x, 16); }

static inline fe fe_add(fe a, fe b) {
    u128 s = a + b;
    /* a,b < M: true sum < 2M < 2^129; wrapped iff s < a */
    if (s < a || s >= FE_M) s -= FE_M;
    return s;
}
static inline fe fe_sub(fe a, fe b) { return a >= b ? a - b : a + (FE_M - b); }
static inline fe fe_neg(fe a) { return a == 0 ? 0 : FE_M - a; }

/* 128x128 -> 256 schoolbook on 64-bit halves, then fold with 2^128 == c */
static inline fe fe_mul(fe a, fe b) {
    uint64_t a0 = (uint64_t)a, a1 = (uint64_t)(a >> 64);
    uint64_t b0 = (uint64_t)b, b1 = (uint64_t)(b >> 64);
    u128 p00 = (u128)a0 * b0, p01 = (u128)a0 * b1, p10 = (u128)a1 * b0, p11 = (u128)a1 * b1;
    u128 mid = (p00 >> 64) + (uint64_t)p01 + (uint64_t)p10;
    u128 lo = ((u128)(uint64_t)mid << 64) | (uint64_t)p00;
    u128 hi = p11 + (p01 >> 64) + (p10 >> 64) + (mid >> 64);
    while (hi != 0) {
        /* hi*c + lo, with hi*c up to 174 bits */
        u128 q0 = (u128)(uint64_t)hi * FE_C;
        u128 q1 = (u128)(uint64_t)(hi >> 64) * FE_C;
        u128 t = q0 + (q1 << 64);
        u128 nhi = (q1 >> 64) + (t < q0);
        u128 s = lo + t;
        nhi += (s < lo);
        lo = s;
        hi = nhi;
    }
    return lo >= FE_M ? lo - FE_M : lo;
}

/* FieldElement::exp — square and multiply (src/ntt/mod.rs:56-58) */
static inline fe fe_pow(fe base, u128 e) {
    fe r = 1;
    while (e) {
        if (e & 1) r = fe_mul(r, base);
        base = fe_mul(base, base);
        e >>= 1;
    }
    return r;
}
static inline fe fe_inv(fe a) { return a == 0 ? 0 : fe_pow(a, FE_M - 2); } /* inv(0)=0 */
static inline fe fe_div(fe a, fe b) { return fe_mul(a, fe_inv(b)); }

#endif
