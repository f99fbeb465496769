/* Goldilocks field helpers for the C oracle.  TEST INFRASTRUCTURE (see oracle/__init__.py).
 * Restates gnark-plonky2-verifier/goldilocks/base.go:33-39 (p = 2^64-2^32+1, g = 7, 2-adic root)
 * and quadratic_extension.go:59-73 (X^2 = 7).  Canonical representation (< p) everywhere. */
#ifndef ORACLE_GL_H
#define ORACLE_GL_H
#include <stdint.h>
#include <stddef.h>
typedef uint64_t u64;
typedef unsigned __int128 u128;
#define GL_P 0xFFFFFFFF00000001ULL
#define GL_GEN 7ULL
#define GL_ROOT32 1753635133440165772ULL

static inline u64 gl_add(u64 a, u64 b) { u128 s = (u128)a + b; if (s >= GL_P) s -= GL_P; return (u64)s; }
static inline u64 gl_sub(u64 a, u64 b) { return a >= b ? a - b : a + (GL_P - b); }
static inline u64 gl_neg(u64 a) { return a ? GL_P - a : 0; }
/* (hi*2^64 + lo) mod p via 2^64 = 2^32 - 1, 2^96 = -1 (the reduction upstream plonky2 uses on the CPU) */
static inline u64 gl_red128(u128 x) {
    u64 lo = (u64)x, hi = (u64)(x >> 64);
    u64 hh = hi >> 32, hl = hi & 0xFFFFFFFFULL;
    u64 t0 = lo - hh; if (lo < hh) t0 -= 0xFFFFFFFFULL;
    u64 t1 = hl * 0xFFFFFFFFULL;
    u64 r = t0 + t1; if (r < t1) r += 0xFFFFFFFFULL;
    return r >= GL_P ? r - GL_P : r;
}
static inline u64 gl_mul(u64 a, u64 b) { return gl_red128((u128)a * b); }
static inline u64 gl_pow(u64 a, u64 e) { u64 r = 1; while (e) { if (e & 1) r = gl_mul(r, a); a = gl_mul(a, a); e >>= 1; } return r; }
static inline u64 gl_inv(u64 a) { return gl_pow(a, GL_P - 2); }
static inline u64 gl_root(int n_log) { u64 r = GL_ROOT32; for (int i = n_log; i < 32; i++) r = gl_mul(r, r); return r; }
static inline u64 bitrev(u64 x, int bits) { u64 r = 0; for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; } return r; }

typedef struct { u64 a, b; } e2;
static inline e2 e2_add(e2 x, e2 y) { e2 r = { gl_add(x.a, y.a), gl_add(x.b, y.b) }; return r; }
static inline e2 e2_sub(e2 x, e2 y) { e2 r = { gl_sub(x.a, y.a), gl_sub(x.b, y.b) }; return r; }
static inline e2 e2_mul(e2 x, e2 y) {
    e2 r = { gl_add(gl_mul(x.a, y.a), gl_mul(7, gl_mul(x.b, y.b))), gl_add(gl_mul(x.a, y.b), gl_mul(x.b, y.a)) };
    return r;
}
static inline e2 e2_scale(e2 x, u64 s) { e2 r = { gl_mul(x.a, s), gl_mul(x.b, s) }; return r; }
static inline e2 e2_inv(e2 x) {
    u64 d = gl_inv(gl_sub(gl_mul(x.a, x.a), gl_mul(7, gl_mul(x.b, x.b))));
    e2 r = { gl_mul(x.a, d), gl_mul(gl_neg(x.b), d) };
    return r;
}
static inline e2 e2_from(u64 a) { e2 r = { a, 0 }; return r; }
static inline int e2_eq(e2 x, e2 y) { return x.a == y.a && x.b == y.b; }
static inline e2 e2_pow(e2 a, u64 e) { e2 r = {1, 0}; while (e) { if (e & 1) r = e2_mul(r, a); a = e2_mul(a, a); e >>= 1; } return r; }
#endif
