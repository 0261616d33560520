// field.cuh -- BabyBear (p = 15*2^27+1) and BabyBear^4 arithmetic for sm_100a.
//
// Reference definitions: basic/src/field/mod.rs:43-64 (MOD = 0x78000001, BinomialExtensionField<_,4>),
// [MEM] p3-baby-bear (Montgomery form, R = 2^32; generator 31; 2^27-th root 0x1a427a41; x^4 = 11).
//
// Two multiplication forms are used on the device:
//   * mmul(a,b)        Montgomery product a*b/R, both operands data (3 multiply-class + 3 ALU instr).
//   * shoup(b,w,w')    b*w mod p for a PRECOMPUTED constant w with w' = floor(w*2^32/p): IMAD.HI + 2 IMAD,
//                      result lazily in [0,2p).  b may be any u32.  Because b carries the Montgomery factor
//                      and w is canonical, the product is again in Montgomery form: NTT data never leaves it.
// Modular add/sub use the unsigned-min trick (IADD, IADD, VIMNMX) instead of compare+select.
#pragma once
#include "ts_platform.h"

namespace bb {

constexpr uint32_t P = 0x78000001u;
constexpr uint32_t PINV = 0x88000001u;       // p^-1 mod 2^32
constexpr uint32_t MONTY_ONE = 0x0ffffffeu;  // R mod p
constexpr uint32_t R2 = 0x45dddde3u;         // R^2 mod p
constexpr uint32_t MONTY_HALF = 0x07ffffffu; // (1/2) * R mod p
constexpr uint32_t MONTY_W = 0x37ffffe9u;    // 11 * R mod p  (x^4 = 11)
constexpr uint32_t ROOT27 = 0x1a427a41u;     // canonical generator of the 2^27 subgroup

TS_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
// High word of a 32 x 32 product.  __umulhi compiles to IMAD.HI, which issues at HALF the rate of the other integer
// multiply-adds on sm_100 (profiles/r01/int_pipes_b200.jsonl: 32 vs 63 lane-ops/clk/SM); a mul.wide whose low half is simply
// not used stays an IMAD.WIDE (full rate) through ptxas.  -DTS_MULHI_WIDE selects it (A/B in profiles/r02/README.md).
TS_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__) && defined(TS_MULHI_WIDE)
    unsigned long long t;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a), "r"(b));
    return (uint32_t)(t >> 32);
#elif defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// c + a * b with a 32 x 32 -> 64 product: ONE IMAD.WIDE.  Spelled as `c + (uint64_t)a * b` it is the same instruction only as long
// as the compiler still knows that the operands are zero-extended 32-bit values; behind predicated loads and selects it forgot
// (opn::bary_partial4_kernel: three IMADs and an IADD3 per product, 41 instructions per element instead of 10).
TS_HD uint64_t madw(uint32_t a, uint32_t b, uint64_t c) {
#if defined(__CUDA_ARCH__)
    unsigned long long t;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(a), "r"(b), "l"((unsigned long long)c));
    return t;
#else
    return c + (uint64_t)a * b;
#endif
}

// inputs in [0,p), output in [0,p)
TS_HD uint32_t add(uint32_t a, uint32_t b) {
    uint32_t s = a + b;
    return umin32(s, s - P);
}
TS_HD uint32_t sub(uint32_t a, uint32_t b) {
    uint32_t d = a - b;
    return umin32(d, d + P);
}
// x in [0,2p) -> [0,p)
TS_HD uint32_t red2p(uint32_t x) { return umin32(x, x - P); }
TS_HD uint32_t neg(uint32_t a) { return a ? P - a : 0; }
// (a/2) mod p
TS_HD uint32_t half(uint32_t a) { return (a >> 1) + ((a & 1) ? (P + 1) / 2 : 0); }

// Montgomery reduction of a 64-bit T < p*2^32: T/R mod p in [0,p)
TS_HD uint32_t redc(uint64_t t) {
    uint32_t lo = (uint32_t)t, hi = (uint32_t)(t >> 32);
    uint32_t m = lo * PINV;
    uint32_t q = mulhi32(m, P);
    uint32_t r = hi - q;
    return umin32(r, r + P);
}
TS_HD uint32_t mmul(uint32_t a, uint32_t b) { return redc((uint64_t)a * b); }
TS_HD uint32_t to_monty(uint32_t x) { return mmul(x, R2); }
TS_HD uint32_t from_monty(uint32_t x) { return redc((uint64_t)x); }

// Shoup / Harvey constant multiplication: result in [0,2p)
TS_HD uint32_t shoup_lazy(uint32_t b, uint32_t w, uint32_t wp) {
    uint32_t q = mulhi32(b, wp);
    return b * w - q * P;
}
TS_HD uint32_t shoup(uint32_t b, uint32_t w, uint32_t wp) { return red2p(shoup_lazy(b, w, wp)); }
TS_HD uint32_t shoup_prime(uint32_t w) { return (uint32_t)((((uint64_t)w) << 32) / P); }

// ---- plain (canonical) helpers for host-side table generation ---------------------------------------
constexpr uint32_t cmul(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) % P); }
constexpr uint32_t cpow(uint32_t a, uint64_t e) {
    uint32_t r = 1;
    while (e) {
        if (e & 1) r = cmul(r, a);
        a = cmul(a, a);
        e >>= 1;
    }
    return r;
}
constexpr uint32_t cinv(uint32_t a) { return cpow(a, (uint64_t)P - 2); }
constexpr uint32_t two_adic_generator(int bits) { return cpow(ROOT27, 1ull << (27 - bits)); }
constexpr uint32_t cshoup_prime(uint32_t w) { return (uint32_t)((((uint64_t)w) << 32) / P); }

// Compile-time twiddles of the in-register radix-2^LOGR DFT: w[j] = root_{2^LOGR}^(+-j), canonical, with
// their Shoup companions.  Fully unrolled uses become immediates.
template <int LOGR, bool INV>
struct InnerTw {
    uint32_t w[(1 << LOGR) / 2 > 0 ? (1 << LOGR) / 2 : 1];
    uint32_t wp[(1 << LOGR) / 2 > 0 ? (1 << LOGR) / 2 : 1];
    constexpr InnerTw() : w{}, wp{} {
        uint32_t root = two_adic_generator(LOGR);
        if (INV) root = cinv(root);
        uint32_t cur = 1;
        for (int j = 0; j < (1 << LOGR) / 2; j++) {
            w[j] = cur;
            wp[j] = cshoup_prime(cur);
            cur = cmul(cur, root);
        }
    }
};

}  // namespace bb

// ---- BabyBear^4, Montgomery-form coefficients -----------------------------------------------------
namespace ef {
struct E4 {
    uint32_t c[4];
};
TS_HD E4 add(const E4 &a, const E4 &b) {
    E4 r;
    for (int i = 0; i < 4; i++) r.c[i] = bb::add(a.c[i], b.c[i]);
    return r;
}
TS_HD E4 sub(const E4 &a, const E4 &b) {
    E4 r;
    for (int i = 0; i < 4; i++) r.c[i] = bb::sub(a.c[i], b.c[i]);
    return r;
}
TS_HD E4 scale(const E4 &a, uint32_t s) {
    E4 r;
    for (int i = 0; i < 4; i++) r.c[i] = bb::mmul(a.c[i], s);
    return r;
}
TS_HD E4 half(const E4 &a) {
    E4 r;
    for (int i = 0; i < 4; i++) r.c[i] = bb::half(a.c[i]);
    return r;
}
// Operand b prepared once: b.c[i] and 11*b.c[i].  a*b over x^4 = 11 with 64-bit accumulation of PAIRS of
// products (2p^2 < p*2^32 keeps redc in range), 2 redc + 1 add per output coefficient.
struct E4Const {
    uint32_t b[4], wb[4];
};
TS_HD E4Const prepare(const E4 &b) {
    E4Const k;
    for (int i = 0; i < 4; i++) {
        k.b[i] = b.c[i];
        k.wb[i] = bb::mmul(b.c[i], bb::MONTY_W);
    }
    return k;
}
TS_HD E4 mul(const E4 &a, const E4Const &k) {
    // c0 = a0b0 + W(a1b3 + a2b2 + a3b1); c1 = a0b1 + a1b0 + W(a2b3 + a3b2)
    // c2 = a0b2 + a1b1 + a2b0 + W a3b3;  c3 = a0b3 + a1b2 + a2b1 + a3b0
    E4 r;
    r.c[0] = bb::add(bb::redc((uint64_t)a.c[0] * k.b[0] + (uint64_t)a.c[1] * k.wb[3]),
                     bb::redc((uint64_t)a.c[2] * k.wb[2] + (uint64_t)a.c[3] * k.wb[1]));
    r.c[1] = bb::add(bb::redc((uint64_t)a.c[0] * k.b[1] + (uint64_t)a.c[1] * k.b[0]),
                     bb::redc((uint64_t)a.c[2] * k.wb[3] + (uint64_t)a.c[3] * k.wb[2]));
    r.c[2] = bb::add(bb::redc((uint64_t)a.c[0] * k.b[2] + (uint64_t)a.c[1] * k.b[1]),
                     bb::redc((uint64_t)a.c[2] * k.b[0] + (uint64_t)a.c[3] * k.wb[3]));
    r.c[3] = bb::add(bb::redc((uint64_t)a.c[0] * k.b[3] + (uint64_t)a.c[1] * k.b[2]),
                     bb::redc((uint64_t)a.c[2] * k.b[1] + (uint64_t)a.c[3] * k.b[0]));
    return r;
}
}  // namespace ef
