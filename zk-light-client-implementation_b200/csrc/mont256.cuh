// 254-bit Montgomery arithmetic on the device for the two BN254 moduli (F_q in msm.cu, F_r in bn254.cuh).
// The 32x32->64 multiplier (IMAD.WIDE, 25 /clk/SM: profiles/int_peak_r01.jsonl) is the scarce unit, so a product spends
// exactly the 128 + 8 wide multiply-accumulates of an 8-limb CIOS and keeps every carry in the flag register:
// T = E + (O << 32); a partial product x_j * y_i goes to E when j + i is even and to O when it is odd, so it always lands on
// an aligned limb pair and a row of four products is ONE carry chain (ptxas fuses each mad.lo.cc / madc.hi.cc pair into
// IMAD.WIDE.U32.X).  Each of the 4 iterations consumes two limbs of b, adds q * m with the 64-bit q that clears the low
// 64 bits, and shifts by 64 bits, which preserves the E/O parity.  250 SASS instructions per product, 136 of them wide
// multiplies (the 64-bit-limb compare/select CIOS this replaces: ~600).  Bounds: E <= T < 2^320 inside an iteration,
// T < 2m after each shift (checked limb for limb against a Python model of the same schedule).
#pragma once
#include "gl.cuh"

#ifdef __CUDACC__
#define MONT_MAC_ROW(c0, c1, c2, c3, c4, c5, c6, c7, c8, x0, x1, x2, x3, y)                                          \
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"  "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"                                    \
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t" "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"                                   \
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t" "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"                                   \
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t" "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"                                   \
        "addc.u32 %8, %8, 0;"                                                                                        \
        : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(c4), "+r"(c5), "+r"(c6), "+r"(c7), "+r"(c8)                  \
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y))
#define MONT_MAC_4ROWS(X, y0, y1)                                                                                    \
    MONT_MAC_ROW(e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7], e[8], X[0], X[2], X[4], X[6], y0);                  \
    MONT_MAC_ROW(o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8], X[1], X[3], X[5], X[7], y0);                  \
    MONT_MAC_ROW(o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8], X[0], X[2], X[4], X[6], y1);                  \
    MONT_MAC_ROW(e[2], e[3], e[4], e[5], e[6], e[7], e[8], e[9], e[10], X[1], X[3], X[5], X[7], y1)

// P: struct with static constexpr u64 M0..M3 (modulus, little-endian limbs) and INV = -m^{-1} mod 2^64.  a, b < m.
template <class P> __device__ __forceinline__ void mont256_mul(const u64 A[4], const u64 B[4], u64 out[4]) {
    const u32 m[8] = {(u32)P::M0, (u32)(P::M0 >> 32), (u32)P::M1, (u32)(P::M1 >> 32), (u32)P::M2, (u32)(P::M2 >> 32), (u32)P::M3, (u32)(P::M3 >> 32)};
    u32 a[8], b[8];
#pragma unroll
    for (int i = 0; i < 4; i++) { a[2 * i] = (u32)A[i]; a[2 * i + 1] = (u32)(A[i] >> 32); b[2 * i] = (u32)B[i]; b[2 * i + 1] = (u32)(B[i] >> 32); }
    u32 e[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, o[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        MONT_MAC_4ROWS(a, b[2 * i], b[2 * i + 1]);
        const u64 q = ((u64)(e[1] + o[0]) << 32 | e[0]) * P::INV;
        const u32 q0 = (u32)q, q1 = (u32)(q >> 32);
        MONT_MAC_4ROWS(m, q0, q1);
        // the low 64 bits of T are now zero: e[1] + o[0] is 0 or 2^32 and its carry belongs to limb 2, as does o[1]; shift by 64
        u32 sink;
        asm("add.cc.u32 %8, %9, %10;\n\t" "addc.cc.u32 %0, %0, %11;\n\t" "addc.cc.u32 %1, %1, 0;\n\t" "addc.cc.u32 %2, %2, 0;\n\t"
            "addc.cc.u32 %3, %3, 0;\n\t" "addc.cc.u32 %4, %4, 0;\n\t" "addc.cc.u32 %5, %5, 0;\n\t" "addc.cc.u32 %6, %6, 0;\n\t" "addc.u32 %7, %7, 0;"
            : "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]), "+r"(e[8]), "+r"(e[9]), "=r"(sink)
            : "r"(e[1]), "r"(o[0]), "r"(o[1]));
        (void)sink;
#pragma unroll
        for (int k = 0; k < 8; k++) e[k] = e[k + 2];
        e[8] = e[9] = e[10] = 0;
#pragma unroll
        for (int k = 0; k < 7; k++) o[k] = o[k + 2];
        o[7] = o[8] = 0;
    }
    // T = E + (O << 32) < 2m, then one conditional subtraction
    u32 r[8]; r[0] = e[0];
    asm("add.cc.u32 %0, %7, %14;\n\t" "addc.cc.u32 %1, %8, %15;\n\t" "addc.cc.u32 %2, %9, %16;\n\t" "addc.cc.u32 %3, %10, %17;\n\t"
        "addc.cc.u32 %4, %11, %18;\n\t" "addc.cc.u32 %5, %12, %19;\n\t" "addc.u32 %6, %13, %20;"
        : "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]));
    u32 t[8], borrow;
    asm("sub.cc.u32 %0, %9, %17;\n\t" "subc.cc.u32 %1, %10, %18;\n\t" "subc.cc.u32 %2, %11, %19;\n\t" "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t" "subc.cc.u32 %5, %14, %22;\n\t" "subc.cc.u32 %6, %15, %23;\n\t" "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(borrow)
        : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(m[0]), "r"(m[1]), "r"(m[2]), "r"(m[3]), "r"(m[4]), "r"(m[5]), "r"(m[6]), "r"(m[7]));
#pragma unroll
    for (int i = 0; i < 4; i++) {
        u32 lo = borrow ? r[2 * i] : t[2 * i], hi = borrow ? r[2 * i + 1] : t[2 * i + 1];
        out[i] = ((u64)hi << 32) | lo;
    }
}
// (a + b) mod m and (a - b) mod m for a, b < m < 2^254: 64-bit carry chains, no compares
template <class P> __device__ __forceinline__ void mont256_add(const u64 a[4], const u64 b[4], u64 out[4]) {
    u64 s0, s1, s2, s3, t0, t1, t2, t3; u32 borrow;
    const u64 m0 = P::M0, m1 = P::M1, m2 = P::M2, m3 = P::M3;
    asm("add.cc.u64 %0, %4, %8;\n\t" "addc.cc.u64 %1, %5, %9;\n\t" "addc.cc.u64 %2, %6, %10;\n\t" "addc.u64 %3, %7, %11;"
        : "=l"(s0), "=l"(s1), "=l"(s2), "=l"(s3) : "l"(a[0]), "l"(a[1]), "l"(a[2]), "l"(a[3]), "l"(b[0]), "l"(b[1]), "l"(b[2]), "l"(b[3]));
    asm("sub.cc.u64 %0, %5, %9;\n\t" "subc.cc.u64 %1, %6, %10;\n\t" "subc.cc.u64 %2, %7, %11;\n\t" "subc.cc.u64 %3, %8, %12;\n\t" "subc.u32 %4, 0, 0;"
        : "=l"(t0), "=l"(t1), "=l"(t2), "=l"(t3), "=r"(borrow) : "l"(s0), "l"(s1), "l"(s2), "l"(s3), "l"(m0), "l"(m1), "l"(m2), "l"(m3));
    out[0] = borrow ? s0 : t0; out[1] = borrow ? s1 : t1; out[2] = borrow ? s2 : t2; out[3] = borrow ? s3 : t3;
}
template <class P> __device__ __forceinline__ void mont256_sub(const u64 a[4], const u64 b[4], u64 out[4]) {
    u64 d0, d1, d2, d3; u32 borrow;
    asm("sub.cc.u64 %0, %5, %9;\n\t" "subc.cc.u64 %1, %6, %10;\n\t" "subc.cc.u64 %2, %7, %11;\n\t" "subc.cc.u64 %3, %8, %12;\n\t" "subc.u32 %4, 0, 0;"
        : "=l"(d0), "=l"(d1), "=l"(d2), "=l"(d3), "=r"(borrow) : "l"(a[0]), "l"(a[1]), "l"(a[2]), "l"(a[3]), "l"(b[0]), "l"(b[1]), "l"(b[2]), "l"(b[3]));
    const u64 k0 = borrow ? P::M0 : 0, k1 = borrow ? P::M1 : 0, k2 = borrow ? P::M2 : 0, k3 = borrow ? P::M3 : 0;
    asm("add.cc.u64 %0, %4, %8;\n\t" "addc.cc.u64 %1, %5, %9;\n\t" "addc.cc.u64 %2, %6, %10;\n\t" "addc.u64 %3, %7, %11;"
        : "=l"(out[0]), "=l"(out[1]), "=l"(out[2]), "=l"(out[3]) : "l"(d0), "l"(d1), "l"(d2), "l"(d3), "l"(k0), "l"(k1), "l"(k2), "l"(k3));
}
#endif
