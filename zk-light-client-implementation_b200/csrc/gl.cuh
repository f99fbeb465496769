// Goldilocks field (p = 2^64 - 2^32 + 1) and its quadratic extension X^2 = 7, host + device.
// Spec: plonky2_field::goldilocks_field (upstream, not vendored) as mirrored by
// gnark-plonky2-verifier/goldilocks/base.go:33-39 and quadratic_extension.go:59-73.
// All values stored in memory are canonical (< p).  Internally `gl_mul_nc` may return a
// non-canonical representative in [0, 2^64); every *_nc result must go through gl_canon
// before it is stored or compared.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t u8;

#define GL_P   0xFFFFFFFF00000001ULL
#define GL_EPS 0x00000000FFFFFFFFULL
#define GL_GEN 7ULL
#define GL_ROOT32 1753635133440165772ULL   // primitive 2^32-th root of unity (base.go:37)

#ifdef __CUDACC__
#define GL_HD __host__ __device__ __forceinline__
#else
#define GL_HD inline
#endif

// Device versions are 32-bit carry chains in PTX: the compiler's translation of the C forms below goes through predicates
// (ISETP / SEL pairs, ~11 instructions per modular add), and these kernels are bound by the integer ALU pipe
// (profiles/ncu_ntt2_r02_v3.jsonl: alu 75 %, fma 35 %): 9 / 5 / 6 instructions for add / sub / canonicalise.
GL_HD u64 gl_canon(u64 a) {          // any u64 -> canonical
#ifdef __CUDA_ARCH__
    u32 r0, r1;
    asm("{\n\t.reg .u32 t0, t1, c;\n\t"
        "add.cc.u32 t0, %2, 0xFFFFFFFF;\n\t"      // carry out of a + EPS  <=>  a >= p
        "addc.cc.u32 t1, %3, 0;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "neg.s32 c, c;\n\t"                        // mask == EPS when a >= p
        "add.cc.u32 %0, %2, c;\n\t"                // a + EPS == a - p (mod 2^64)
        "addc.u32 %1, %3, 0;\n\t}"
        : "=&r"(r0), "=&r"(r1) : "r"((u32)a), "r"((u32)(a >> 32)));
    return ((u64)r1 << 32) | r0;
#else
    return a >= GL_P ? a - GL_P : a;
#endif
}

// a, b canonical -> canonical
GL_HD u64 gl_add(u64 a, u64 b) {
#ifdef __CUDA_ARCH__
    u32 r0, r1;
    asm("{\n\t.reg .u32 s0, s1, t0, t1, c;\n\t"
        "add.cc.u32 s0, %2, %4;\n\t"
        "addc.cc.u32 s1, %3, %5;\n\t"
        "addc.u32 c, 0, 0;\n\t"                    // a + b >= 2^64
        "add.cc.u32 t0, s0, 0xFFFFFFFF;\n\t"       // or (a + b mod 2^64) >= p: never both
        "addc.cc.u32 t1, s1, 0;\n\t"
        "addc.u32 c, c, 0;\n\t"
        "neg.s32 c, c;\n\t"
        "add.cc.u32 %0, s0, c;\n\t"                // subtract p in either case (== add EPS mod 2^64)
        "addc.u32 %1, s1, 0;\n\t}"
        : "=&r"(r0), "=&r"(r1) : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
    return ((u64)r1 << 32) | r0;
#else
    u64 s = a + b;
    if (s < a) s += GL_EPS;            // wrapped past 2^64: 2^64 == EPS (mod p); result already < p
    else if (s >= GL_P) s -= GL_P;
    return s;
#endif
}
// a: ANY 64-bit representative, b: canonical -> some representative of a + b in [0, 2^64) (6 instructions instead of 9:
// with b < p the wrapped sum plus EPS cannot wrap a second time).  For sums that feed a multiplication.
GL_HD u64 gl_add_nc1(u64 a, u64 b) {
#ifdef __CUDA_ARCH__
    u32 r0, r1;
    asm("{\n\t.reg .u32 c;\n\t"
        "add.cc.u32 %0, %2, %4;\n\t"
        "addc.cc.u32 %1, %3, %5;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "neg.s32 c, c;\n\t"                        // EPS when a + b >= 2^64
        "add.cc.u32 %0, %0, c;\n\t"
        "addc.u32 %1, %1, 0;\n\t}"
        : "=&r"(r0), "=&r"(r1) : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
    return ((u64)r1 << 32) | r0;
#else
    u64 s = a + b;
    if (s < a) s += GL_EPS;
    return s;
#endif
}
GL_HD u64 gl_sub(u64 a, u64 b) {
#ifdef __CUDA_ARCH__
    u32 r0, r1;
    asm("{\n\t.reg .u32 m;\n\t"
        "sub.cc.u32 %0, %2, %4;\n\t"
        "subc.cc.u32 %1, %3, %5;\n\t"
        "subc.u32 m, 0, 0;\n\t"                    // borrow mask == EPS
        "sub.cc.u32 %0, %0, m;\n\t"                // + p  ==  - EPS (mod 2^64)
        "subc.u32 %1, %1, 0;\n\t}"
        : "=&r"(r0), "=&r"(r1) : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
    return ((u64)r1 << 32) | r0;
#else
    u64 d = a - b;
    if (a < b) d -= GL_EPS;            // + p  ==  - EPS (mod 2^64)
    return d;
#endif
}
GL_HD u64 gl_neg(u64 a) { return a ? GL_P - a : 0; }

// x * 2^T mod p for a compile-time 0 < T < 96 with T % 32 != 0, x canonical -> canonical.  2 has order 192 in F_p (2^96 == -1), so
// the 16th / 64th roots of unity are powers of two and small DFTs need no multiplier: with phi = 2^32 (phi^2 = phi - 1,
// phi^3 = -1) the shifted value w_lo + w_mid phi + w_hi phi^2 times phi^(T / 32) collapses to two-word sums.
template <int T> GL_HD u64 gl_mul_pow2(u64 x) {
    static_assert(T > 0 && T < 96 && (T % 32) != 0, "shift out of range");
    constexpr int t = T % 32, q = T / 32;
    const u32 a = (u32)x, b = (u32)(x >> 32);
    const u32 w_lo = a << t, w_mid = (u32)(x >> (32 - t)), w_hi = b >> (32 - t);
    if (q == 0) {                   // w_lo + w_mid phi + w_hi (phi - 1)
        const u64 X = ((u64)w_mid << 32) | w_lo, t1 = ((u64)w_hi << 32) - w_hi;
        return gl_canon(gl_add_nc1(X, t1));
    }
    if (q == 1) {                   // w_lo phi + w_mid (phi - 1) - w_hi
        const u64 X = (u64)w_lo << 32, t1 = ((u64)w_mid << 32) - w_mid;        // both < p
        return gl_add(X, gl_sub(t1, (u64)w_hi));
    }
    // q == 2:                         w_lo (phi - 1) - (w_mid + w_hi phi)
    const u64 t1 = ((u64)w_lo << 32) - w_lo, Y = ((u64)w_hi << 32) | w_mid;    // both < p
    return gl_sub(t1, Y);
}

GL_HD void gl_mul_wide(u64 a, u64 b, u64 &lo, u64 &hi) {
#ifdef __CUDA_ARCH__
    // Four INDEPENDENT IMAD.WIDE.U32 and a six-instruction carry chain.  (A chained form -- each product added onto the
    // shifted previous one -- needs its 64-bit addend pair rebuilt before every IMAD.WIDE: two extra moves per product in
    // SASS.  The 32x32->64 multiplier is the scarce unit on sm_100a: 25 /clk/SM, profiles/int_peak_r01.jsonl.)
    u32 x0, x1, x2, x3;
    asm("{\n\t"
        ".reg .u64 p00, p01, p10, p11;\n\t"
        ".reg .u32 a0, a1, b0, b1, l01, h01, l10, h10, l11, h11, h00, m0, m1, m2;\n\t"
        "mov.b64 {a0, a1}, %4;\n\t"
        "mov.b64 {b0, b1}, %5;\n\t"
        "mul.wide.u32 p00, a0, b0;\n\t"
        "mul.wide.u32 p01, a0, b1;\n\t"
        "mul.wide.u32 p10, a1, b0;\n\t"
        "mul.wide.u32 p11, a1, b1;\n\t"
        "mov.b64 {%0, h00}, p00;\n\t"
        "mov.b64 {l01, h01}, p01;\n\t"
        "mov.b64 {l10, h10}, p10;\n\t"
        "mov.b64 {l11, h11}, p11;\n\t"
        "add.cc.u32 m0, l01, l10;\n\t"          // mid = p01 + p10 (65 bits: m2:m1:m0)
        "addc.cc.u32 m1, h01, h10;\n\t"
        "addc.u32 m2, 0, 0;\n\t"
        "add.cc.u32 %1, h00, m0;\n\t"
        "addc.cc.u32 %2, l11, m1;\n\t"
        "addc.u32 %3, h11, m2;\n\t"
        "}" : "=r"(x0), "=&r"(x1), "=&r"(x2), "=&r"(x3) : "l"(a), "l"(b));
    lo = ((u64)x1 << 32) | x0; hi = ((u64)x3 << 32) | x2;
#else
    unsigned __int128 p = (unsigned __int128)a * b; lo = (u64)p; hi = (u64)(p >> 64);
#endif
}

// (hi*2^64 + lo) mod p, result in [0, 2^64) (not necessarily canonical).
// 2^64 == 2^32 - 1, 2^96 == -1 (mod p).
GL_HD u64 gl_reduce128_nc(u64 lo, u64 hi) {
#ifdef __CUDA_ARCH__
    u32 rl, rh;
    asm("{\n\t"
        ".reg .u32 t0l, t0h, bm, t1l, t1h, cm;\n\t"
        "sub.cc.u32 t0l, %2, %5;\n\t"        // t0 = lo - (hi >> 32)
        "subc.cc.u32 t0h, %3, 0;\n\t"
        "subc.u32 bm, 0, 0;\n\t"             // borrow mask: 0 or 0xFFFFFFFF (== EPS)
        "sub.cc.u32 t0l, t0l, bm;\n\t"       // t0 -= EPS on borrow
        "subc.u32 t0h, t0h, 0;\n\t"
        "sub.cc.u32 t1l, 0, %4;\n\t"         // t1 = (hi_lo << 32) - hi_lo
        "subc.u32 t1h, %4, 0;\n\t"
        "add.cc.u32 %0, t0l, t1l;\n\t"
        "addc.cc.u32 %1, t0h, t1h;\n\t"
        "addc.u32 cm, 0, 0;\n\t"
        "neg.s32 cm, cm;\n\t"                // carry mask
        "add.cc.u32 %0, %0, cm;\n\t"         // += EPS on carry
        "addc.u32 %1, %1, 0;\n\t"
        "}" : "=&r"(rl), "=&r"(rh) : "r"((u32)lo), "r"((u32)(lo >> 32)), "r"((u32)hi), "r"((u32)(hi >> 32)));
    return ((u64)rh << 32) | rl;
#else
    u64 hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    u64 t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= GL_EPS;
    u64 t1 = hi_lo * GL_EPS;           // (hi_lo << 32) - hi_lo, fits in 64 bits
    u64 r = t0 + t1;
    if (r < t1) r += GL_EPS;
    return r;
#endif
}
GL_HD u64 gl_mul_nc(u64 a, u64 b) { u64 lo, hi; gl_mul_wide(a, b, lo, hi); return gl_reduce128_nc(lo, hi); }
GL_HD u64 gl_mul(u64 a, u64 b) { return gl_canon(gl_mul_nc(a, b)); }
GL_HD u64 gl_sqr(u64 a) { return gl_mul(a, a); }

GL_HD u64 gl_pow(u64 a, u64 e) {
    u64 r = 1;
    while (e) { if (e & 1) r = gl_mul(r, a); a = gl_mul(a, a); e >>= 1; }
    return r;
}
GL_HD u64 gl_inv(u64 a) { return gl_pow(a, GL_P - 2); }
GL_HD u64 gl_root_of_unity(int n_log) {       // base.go:445-455
    u64 r = GL_ROOT32;
    for (int i = n_log; i < 32; i++) r = gl_mul(r, r);
    return r;
}
GL_HD u32 bitrev32(u32 x, int bits) {
#ifdef __CUDA_ARCH__
    return bits ? (__brev(x) >> (32 - bits)) : 0;
#else
    u32 r = 0; for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; } return r;
#endif
}

// ---- quadratic extension ----
struct gl2 { u64 a, b; };
GL_HD gl2 gl2_make(u64 a, u64 b) { gl2 r; r.a = a; r.b = b; return r; }
GL_HD gl2 gl2_add(gl2 x, gl2 y) { return gl2_make(gl_add(x.a, y.a), gl_add(x.b, y.b)); }
GL_HD gl2 gl2_sub(gl2 x, gl2 y) { return gl2_make(gl_sub(x.a, y.a), gl_sub(x.b, y.b)); }
GL_HD gl2 gl2_mul(gl2 x, gl2 y) {
    u64 bb = gl_mul(x.b, y.b);
    u64 w = gl_add(gl_add(gl_add(bb, bb), gl_add(bb, bb)), gl_add(gl_add(bb, bb), bb));   // 7*bb
    return gl2_make(gl_add(gl_mul(x.a, y.a), w), gl_add(gl_mul(x.a, y.b), gl_mul(x.b, y.a)));
}
GL_HD gl2 gl2_scale(gl2 x, u64 s) { return gl2_make(gl_mul(x.a, s), gl_mul(x.b, s)); }
GL_HD gl2 gl2_add_base(gl2 x, u64 s) { return gl2_make(gl_add(x.a, s), x.b); }
GL_HD gl2 gl2_inv(gl2 x) {
    u64 d = gl_inv(gl_sub(gl_mul(x.a, x.a), gl_mul(7, gl_mul(x.b, x.b))));
    return gl2_make(gl_mul(x.a, d), gl_mul(gl_neg(x.b), d));
}
GL_HD gl2 gl2_pow(gl2 a, u64 e) {
    gl2 r = gl2_make(1, 0);
    while (e) { if (e & 1) r = gl2_mul(r, a); a = gl2_mul(a, a); e >>= 1; }
    return r;
}
GL_HD bool gl2_eq(gl2 x, gl2 y) { return x.a == y.a && x.b == y.b; }
