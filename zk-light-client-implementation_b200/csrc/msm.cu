// BN254 G1 multi-scalar multiplication (Pippenger bucket method) for sm_100a -- SURVEY.md section 8f rank 3 / BASELINE
// config 4: the MSMs inside gnark's groth16.Prove (gnark-plonky2-verifier/cmd/web-api.go:77; gnark-crypto
// v0.12.2-0.20231013160410-1f65e75b6dfb ecc/bn254 MultiExp, not vendored) that wrap the final plonky2 proof.
// Result = sum_i scalars[i] * points[i] on y^2 = x^3 + 3 over F_q, returned in affine coordinates: a group element, so
// bit-exact against any correct implementation (oracle/bn254_curve.py restates the textbook group law).
//
// Layout: points affine (x, y) = 8 u64 little-endian limbs, canonical (NOT Montgomery) -- the (0,0) encoding is the
// point at infinity, as in gnark-crypto; scalars 4 u64 limbs canonical < r.  Window c = 16 bits, 16 windows.
//   1. k_digits      : one thread per scalar, 16 digits, histogram per (window, bucket)               (reads 32 N bytes)
//   2. k_scan        : exclusive scan of the 65536 bucket counts of each window
//   3. k_scatter     : counting-sort the point indices by bucket
//   4. k_buckets     : one thread per (window, bucket): XYZZ accumulator += affine point (8M + 2S)  (reads 64 B per digit)
//   5. k_window_*    : per window sum_d d * B_d via chunked running sums, then Horner over the windows (host, 16 steps)
// Algorithmic bytes: 64 N (points) + 32 N (scalars); the bucket phase re-reads every point 16 times through L2.
#include "../../include/b200_plonky2.h"
#include "ctx.h"
#include "mont256.cuh"
#include <vector>

namespace {

struct fq { u64 v[4]; };
#define Q0 0x3c208c16d87cfd47ULL
#define Q1 0x97816a916871ca8dULL
#define Q2 0xb85045b68181585dULL
#define Q3 0x30644e72e131a029ULL
#define Q_INV 0x87d20782e4866389ULL

struct FqMod { static constexpr u64 M0 = Q0, M1 = Q1, M2 = Q2, M3 = Q3, INV = Q_INV; };
#ifndef __CUDA_ARCH__       // host CIOS only; the device product is mont256_mul
GL_HD u64 mac64(u64 a, u64 b, u64 c, u64 &carry) {
    u64 lo, hi; gl_mul_wide(a, b, lo, hi);
    lo += c; hi += (lo < c);
    lo += carry; hi += (lo < carry);
    carry = hi; return lo;
}
#endif
GL_HD bool fq_geq(const u64 a[4]) {
    const u64 m[4] = {Q0, Q1, Q2, Q3};
    for (int i = 3; i >= 0; i--) { if (a[i] > m[i]) return true; if (a[i] < m[i]) return false; }
    return true;
}
GL_HD void fq_subq(u64 a[4]) {
    const u64 m[4] = {Q0, Q1, Q2, Q3}; u64 br = 0;
    for (int i = 0; i < 4; i++) { u64 d = a[i] - m[i]; u64 b1 = a[i] < m[i]; u64 d2 = d - br; u64 b2 = d < br; a[i] = d2; br = b1 | b2; }
}
GL_HD fq fq_add(const fq &a, const fq &b) {
#ifdef __CUDA_ARCH__
    fq rr; mont256_add<FqMod>(a.v, b.v, rr.v); return rr;
#endif
    fq r; u64 c = 0;
    for (int i = 0; i < 4; i++) { u64 s = a.v[i] + b.v[i]; u64 c1 = s < a.v[i]; u64 s2 = s + c; u64 c2 = s2 < s; r.v[i] = s2; c = c1 | c2; }
    if (c || fq_geq(r.v)) fq_subq(r.v);
    return r;
}
GL_HD fq fq_sub(const fq &a, const fq &b) {
#ifdef __CUDA_ARCH__
    fq rr; mont256_sub<FqMod>(a.v, b.v, rr.v); return rr;
#endif
    const u64 m[4] = {Q0, Q1, Q2, Q3};
    fq r; u64 br = 0;
    for (int i = 0; i < 4; i++) { u64 d = a.v[i] - b.v[i]; u64 b1 = a.v[i] < b.v[i]; u64 d2 = d - br; u64 b2 = d < br; r.v[i] = d2; br = b1 | b2; }
    if (br) { u64 c = 0; for (int i = 0; i < 4; i++) { u64 s = r.v[i] + m[i]; u64 c1 = s < r.v[i]; u64 s2 = s + c; u64 c2 = s2 < s; r.v[i] = s2; c = c1 | c2; } }
    return r;
}
#define FQ_FN static __host__ __device__ __noinline__
FQ_FN fq fq_mul(const fq &a, const fq &b) {
#ifdef __CUDA_ARCH__
    fq r; mont256_mul<FqMod>(a.v, b.v, r.v); return r;          // 32-bit-limb even/odd CIOS, mont256.cuh
#else
    const u64 m[4] = {Q0, Q1, Q2, Q3};                           // host: Montgomery CIOS on 64-bit limbs
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u64 c = 0;
        for (int j = 0; j < 4; j++) t[j] = mac64(a.v[j], b.v[i], t[j], c);
        { u64 s = t[4] + c; t[5] = s < c; t[4] = s; }
        u64 qq = t[0] * Q_INV; c = 0;
        (void)mac64(qq, m[0], t[0], c);
        for (int j = 1; j < 4; j++) t[j - 1] = mac64(qq, m[j], t[j], c);
        { u64 s = t[4] + c; u64 cc = s < c; t[3] = s; t[4] = t[5] + cc; }
    }
    fq r; for (int i = 0; i < 4; i++) r.v[i] = t[i];
    if (t[4] || fq_geq(r.v)) fq_subq(r.v);
    return r;
#endif
}
GL_HD fq fq_zero() { fq r; for (int i = 0; i < 4; i++) r.v[i] = 0; return r; }
GL_HD bool fq_is_zero(const fq &a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
GL_HD fq fq_to_mont(const u64 x[4]) {
    const fq r2 = {{0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}};
    fq a; for (int i = 0; i < 4; i++) a.v[i] = x[i];
    return fq_mul(a, r2);
}
GL_HD void fq_from_mont(const fq &a, u64 out[4]) { const fq one = {{1, 0, 0, 0}}; fq r = fq_mul(a, one); for (int i = 0; i < 4; i++) out[i] = r.v[i]; }
GL_HD fq fq_one() { const fq o = {{0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL}}; return o; }

static fq fq_inv_host(const fq &a) {      // Fermat, host only (final conversion to affine)
    u64 e[4] = {Q0 - 2, Q1, Q2, Q3};
    fq r = fq_one(), b = a;
    for (int i = 0; i < 256; i++) { if ((e[i / 64] >> (i % 64)) & 1) r = fq_mul(r, b); b = fq_mul(b, b); }
    return r;
}

// F_q2 = F_q[u]/(u^2 + 1) for G2 (the twist y^2 = x^3 + 3/(9+u)); a = a0 + a1 u
struct fq2 { fq a, b; };
// one spelling of the field operations for both base fields, so the group law and the kernels are written once
GL_HD fq f_add(const fq &x, const fq &y) { return fq_add(x, y); }
GL_HD fq f_sub(const fq &x, const fq &y) { return fq_sub(x, y); }
GL_HD fq f_mul(const fq &x, const fq &y) { return fq_mul(x, y); }
GL_HD fq f_sqr(const fq &x) { return fq_mul(x, x); }
GL_HD bool f_is_zero(const fq &x) { return fq_is_zero(x); }
GL_HD void f_set_zero(fq &x) { x = fq_zero(); }
GL_HD void f_set_one(fq &x) { x = fq_one(); }
GL_HD fq2 f_add(const fq2 &x, const fq2 &y) { fq2 r; r.a = fq_add(x.a, y.a); r.b = fq_add(x.b, y.b); return r; }
GL_HD fq2 f_sub(const fq2 &x, const fq2 &y) { fq2 r; r.a = fq_sub(x.a, y.a); r.b = fq_sub(x.b, y.b); return r; }
GL_HD fq2 f_mul(const fq2 &x, const fq2 &y) {        // Karatsuba: 3 base multiplications
    fq t0 = fq_mul(x.a, y.a), t1 = fq_mul(x.b, y.b), t2 = fq_mul(fq_add(x.a, x.b), fq_add(y.a, y.b));
    fq2 r; r.a = fq_sub(t0, t1); r.b = fq_sub(fq_sub(t2, t0), t1); return r;
}
GL_HD fq2 f_sqr(const fq2 &x) {                       // (a+b)(a-b) + 2ab u
    fq t = fq_mul(x.a, x.b);
    fq2 r; r.a = fq_mul(fq_add(x.a, x.b), fq_sub(x.a, x.b)); r.b = fq_add(t, t); return r;
}
GL_HD bool f_is_zero(const fq2 &x) { return fq_is_zero(x.a) && fq_is_zero(x.b); }
GL_HD void f_set_zero(fq2 &x) { x.a = fq_zero(); x.b = fq_zero(); }
GL_HD void f_set_one(fq2 &x) { x.a = fq_one(); x.b = fq_zero(); }
static fq f_inv_host(const fq &x) { return fq_inv_host(x); }
static fq2 f_inv_host(const fq2 &x) {
    fq d = fq_inv_host(fq_add(fq_mul(x.a, x.a), fq_mul(x.b, x.b)));
    fq2 r; r.a = fq_mul(x.a, d); r.b = fq_sub(fq_zero(), fq_mul(x.b, d)); return r;
}
// affine coordinates <-> canonical little-endian limbs (4 per F_q element; F_q2 as a0 then a1, gnark-crypto's E2{A0, A1})
GL_HD void f_store(const fq &x, u64 *p) { fq_from_mont(x, p); }
GL_HD void f_store(const fq2 &x, u64 *p) { fq_from_mont(x.a, p); fq_from_mont(x.b, p + 4); }
template <class F> struct limbs_of { static constexpr int N = sizeof(F) / 8; };

template <class F> struct xyzz_t { F x, y, zz, zzz; };      // affine (x/zz, y/zzz); identity: zz == 0
template <class F> GL_HD xyzz_t<F> pt_identity() { xyzz_t<F> p; f_set_zero(p.x); f_set_zero(p.y); f_set_zero(p.zz); f_set_zero(p.zzz); return p; }
template <class F> GL_HD bool pt_is_identity(const xyzz_t<F> &p) { return f_is_zero(p.zz); }
#define PT_FN template <class F> static __host__ __device__ __noinline__

PT_FN xyzz_t<F> pt_double_affine(const F &x, const F &y) {     // dbl-2008-s-1 with ZZ = ZZZ = 1, a = 0
    F u = f_add(y, y), v = f_sqr(u), w = f_mul(u, v), s = f_mul(x, v);
    F xx = f_sqr(x), m = f_add(f_add(xx, xx), xx);
    xyzz_t<F> r;
    r.x = f_sub(f_sqr(m), f_add(s, s));
    r.y = f_sub(f_mul(m, f_sub(s, r.x)), f_mul(w, y));
    r.zz = v; r.zzz = w;
    return r;
}
PT_FN xyzz_t<F> pt_double(const xyzz_t<F> &p) {
    if (pt_is_identity(p)) return p;
    F u = f_add(p.y, p.y), v = f_sqr(u), w = f_mul(u, v), s = f_mul(p.x, v);
    F xx = f_sqr(p.x), m = f_add(f_add(xx, xx), xx);
    xyzz_t<F> r;
    r.x = f_sub(f_sqr(m), f_add(s, s));
    r.y = f_sub(f_mul(m, f_sub(s, r.x)), f_mul(w, p.y));
    r.zz = f_mul(v, p.zz); r.zzz = f_mul(w, p.zzz);
    return r;
}
PT_FN void pt_madd(xyzz_t<F> &a, const F &x2, const F &y2) {   // a += (x2, y2) affine (madd-2008-s)
    if (pt_is_identity(a)) { a.x = x2; a.y = y2; f_set_one(a.zz); f_set_one(a.zzz); return; }
    F u2 = f_mul(x2, a.zz), s2 = f_mul(y2, a.zzz);
    F p = f_sub(u2, a.x), r = f_sub(s2, a.y);
    if (f_is_zero(p)) {
        if (f_is_zero(r)) a = pt_double_affine(x2, y2); else a = pt_identity<F>();
        return;
    }
    F pp = f_sqr(p), ppp = f_mul(p, pp), q = f_mul(a.x, pp);
    F x3 = f_sub(f_sub(f_sqr(r), ppp), f_add(q, q));
    F y3 = f_sub(f_mul(r, f_sub(q, x3)), f_mul(a.y, ppp));
    a.x = x3; a.y = y3; a.zz = f_mul(a.zz, pp); a.zzz = f_mul(a.zzz, ppp);
}
PT_FN void pt_add(xyzz_t<F> &a, const xyzz_t<F> &b) {          // a += b (add-2008-s)
    if (pt_is_identity(b)) return;
    if (pt_is_identity(a)) { a = b; return; }
    F u1 = f_mul(a.x, b.zz), u2 = f_mul(b.x, a.zz), s1 = f_mul(a.y, b.zzz), s2 = f_mul(b.y, a.zzz);
    F p = f_sub(u2, u1), r = f_sub(s2, s1);
    if (f_is_zero(p)) {
        if (f_is_zero(r)) a = pt_double(a); else a = pt_identity<F>();
        return;
    }
    F pp = f_sqr(p), ppp = f_mul(p, pp), q = f_mul(u1, pp);
    F x3 = f_sub(f_sub(f_sqr(r), ppp), f_add(q, q));
    F y3 = f_sub(f_mul(r, f_sub(q, x3)), f_mul(s1, ppp));
    a.x = x3; a.y = y3; a.zz = f_mul(f_mul(a.zz, b.zz), pp); a.zzz = f_mul(f_mul(a.zzz, b.zzz), ppp);
}

constexpr int WBITS = 16, NWIN = 16;            // 16 x 16 = 256 >= 254 scalar bits
constexpr u32 NBUCKET = 1u << WBITS;

__global__ void k_digits(const u64 *__restrict__ scalars, size_t n, u32 *__restrict__ digits, u32 *__restrict__ counts) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        u64 w = scalars[4 * i + k];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            u32 d = (u32)(w >> (16 * j)) & 0xFFFFu;
            int win = 4 * k + j;
            digits[(size_t)win * n + i] = d;
            if (d) atomicAdd(&counts[(size_t)win * NBUCKET + d], 1u);
        }
    }
}
__global__ void __launch_bounds__(1024) k_scan(const u32 *__restrict__ counts, u32 *__restrict__ offsets) {      // one block per window, 1024 threads x 64 buckets
    __shared__ u32 part[1024];
    const u32 *c = counts + (size_t)blockIdx.x * NBUCKET; u32 *o = offsets + (size_t)blockIdx.x * NBUCKET;
    u32 base = threadIdx.x * 64, s = 0;
    for (int k = 0; k < 64; k++) s += c[base + k];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) { u32 t = threadIdx.x >= off ? part[threadIdx.x - off] : 0; __syncthreads(); part[threadIdx.x] += t; __syncthreads(); }
    u32 run = threadIdx.x ? part[threadIdx.x - 1] : 0;
    for (int k = 0; k < 64; k++) { o[base + k] = run; run += c[base + k]; }
}
__global__ void k_scatter(const u32 *__restrict__ digits, size_t n, const u32 *__restrict__ offsets, u32 *__restrict__ cursor, u32 *__restrict__ sorted) {
    size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= n * NWIN) return;
    size_t win = idx / n, i = idx % n;
    u32 d = digits[idx];
    if (!d) return;
    u32 pos = atomicAdd(&cursor[win * NBUCKET + d], 1u);
    sorted[win * n + offsets[win * NBUCKET + d] + pos] = (u32)i;
}
// affine points, canonical limbs -> Montgomery limbs, once per MSM (every point is added into 16 buckets); the all-zero
// encoding of the point at infinity is its own image
__global__ void k_points_to_mont(const u64 *__restrict__ in, u64 *__restrict__ out, size_t n_fq) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n_fq) return;
    u64 l[4] = {in[4 * i], in[4 * i + 1], in[4 * i + 2], in[4 * i + 3]};
    fq x = fq_to_mont(l);
#pragma unroll
    for (int k = 0; k < 4; k++) out[4 * i + k] = x.v[k];
}
GL_HD void f_load_mont(fq &x, const u64 *p) { for (int i = 0; i < 4; i++) x.v[i] = p[i]; }
GL_HD void f_load_mont(fq2 &x, const u64 *p) { f_load_mont(x.a, p); f_load_mont(x.b, p + 4); }

template <class F>
__global__ void __launch_bounds__(128) k_buckets(const u64 *__restrict__ points, size_t n, const u32 *__restrict__ counts, const u32 *__restrict__ offsets,
                                                 const u32 *__restrict__ sorted, xyzz_t<F> *__restrict__ buckets) {
    constexpr int NL = limbs_of<F>::N;
    size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (b >= (size_t)NWIN * NBUCKET) return;
    size_t win = b / NBUCKET;
    xyzz_t<F> acc = pt_identity<F>();
    u32 cnt = counts[b], off = offsets[b];
    const u32 *idx = sorted + win * n + off;
    for (u32 k = 0; k < cnt; k++) {
        const u64 *p = points + 2 * NL * (size_t)idx[k];
        u64 l[2 * NL], any = 0;
#pragma unroll
        for (int i = 0; i < 2 * NL; i++) { l[i] = p[i]; any |= l[i]; }
        if (any == 0) continue;                                                                 // point at infinity
        F x, y; f_load_mont(x, l); f_load_mont(y, l + NL);
        pt_madd(acc, x, y);
    }
    buckets[b] = acc;
}
// chunked running sums: thread (win, chunk) covers buckets [start, start + CH)
constexpr u32 CH = 64, NCHUNK = NBUCKET / CH;
template <class F>
__global__ void __launch_bounds__(64) k_window_chunks(const xyzz_t<F> *__restrict__ buckets, xyzz_t<F> *__restrict__ chunk_acc) {
    typedef xyzz_t<F> xyzz;
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= NWIN * NCHUNK) return;
    u32 win = t / NCHUNK, ch = t % NCHUNK, start = ch * CH;
    const xyzz *B = buckets + (size_t)win * NBUCKET;
    xyzz run = pt_identity<F>(), acc = pt_identity<F>();
    for (int d = (int)(start + CH - 1); d >= (int)start; d--) {
        if (d != 0) pt_add(run, B[d]);
        pt_add(acc, run);
    }
    // acc = sum_d (d - start + 1) B_d ; wanted sum_d d B_d = acc + (start - 1) * run
    if (start > 0) {
        u32 k = start - 1;
        xyzz kr = pt_identity<F>(), base = run;
        while (k) { if (k & 1) pt_add(kr, base); base = pt_double(base); k >>= 1; }
        pt_add(acc, kr);
    } else {
        xyzz neg = run; F z; f_set_zero(z); neg.y = f_sub(z, run.y);
        pt_add(acc, neg);
    }
    chunk_acc[t] = acc;
}
// one block of 256 threads per window: each thread sums NCHUNK/256 chunk results, then a shared-memory tree
template <class F>
__global__ void __launch_bounds__(128) k_window_final(const xyzz_t<F> *__restrict__ chunk_acc, xyzz_t<F> *__restrict__ out) {
    typedef xyzz_t<F> xyzz;
    __shared__ xyzz sh[128];                            // 16 KB (G1) / 32 KB (G2)
    const u32 win = blockIdx.x, t = threadIdx.x;
    xyzz s = pt_identity<F>();
    for (u32 c = t; c < NCHUNK; c += 128) pt_add(s, chunk_acc[win * NCHUNK + c]);
    sh[t] = s;
    __syncthreads();
    for (int off = 64; off > 0; off >>= 1) {
        if ((int)t < off) { xyzz a = sh[t]; pt_add(a, sh[t + off]); sh[t] = a; }
        __syncthreads();
    }
    if (t == 0) out[win] = sh[0];
}

}  // namespace

int ctx_fail(b200_ctx *ctx, int code, const std::string &msg);

namespace {
template <class F>
static int msm_dev(b200_ctx *ctx, const uint64_t *d_points, const uint64_t *d_scalars, size_t n, uint64_t *out_affine, const char *who) {
    typedef xyzz_t<F> xyzz;
    constexpr int NL = limbs_of<F>::N;
    if (!ctx) return B200_ERR_BAD_ARG;
    try {
        cudaSetDevice(ctx->device);
        if (!d_points || !d_scalars || !out_affine || n == 0 || n > (size_t(1) << 28)) throw ApiErr{B200_ERR_BAD_ARG, std::string(who) + ": bad arguments"};
        cudaStream_t st = ctx->stream;
        // all scratch comes from the context arena (cudaMalloc/cudaFree per call cost more than the kernels at 2^18 points)
        auto rnd = [](size_t b) { return (b + 255) & ~size_t(255); };
        const size_t sz_digits = rnd(n * NWIN * 4), sz_cnt = rnd((size_t)NWIN * NBUCKET * 4), sz_sorted = rnd(n * NWIN * 4),
                     sz_buckets = rnd((size_t)NWIN * NBUCKET * sizeof(xyzz)), sz_chunks = rnd((size_t)NWIN * NCHUNK * sizeof(xyzz)), sz_wins = rnd(NWIN * sizeof(xyzz)),
                     sz_mont = rnd(n * 2 * NL * 8);
        u8 *base = ctx_arena(ctx, sz_digits + 3 * sz_cnt + sz_sorted + sz_buckets + sz_chunks + sz_wins + sz_mont);
        struct Buf { void *p; size_t bytes; };
        size_t off = 0; auto take = [&](size_t b) { Buf r{base + off, b}; off += b; return r; };
        Buf digits = take(sz_digits), counts = take(sz_cnt), offsets = take(sz_cnt), cursor = take(sz_cnt), sorted = take(sz_sorted),
            buckets = take(sz_buckets), chunks = take(sz_chunks), wins = take(sz_wins), mont = take(sz_mont);
        CUDA_TRY(cudaMemsetAsync(counts.p, 0, counts.bytes, st)); CUDA_TRY(cudaMemsetAsync(cursor.p, 0, cursor.bytes, st));
        { PhaseTimer t(ctx, "msm_to_mont"); k_points_to_mont<<<(unsigned)((n * (NL / 2) + 255) / 256), 256, 0, st>>>(d_points, (u64 *)mont.p, n * (NL / 2)); LAUNCHED(1); }
        { PhaseTimer t(ctx, "msm_digits"); k_digits<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_scalars, n, (u32 *)digits.p, (u32 *)counts.p); LAUNCHED(1); }
        { PhaseTimer t(ctx, "msm_sort");
          k_scan<<<NWIN, 1024, 0, st>>>((const u32 *)counts.p, (u32 *)offsets.p); LAUNCHED(1);
          k_scatter<<<(unsigned)((n * NWIN + 255) / 256), 256, 0, st>>>((const u32 *)digits.p, n, (const u32 *)offsets.p, (u32 *)cursor.p, (u32 *)sorted.p); LAUNCHED(1); }
        { PhaseTimer t(ctx, "msm_buckets"); k_buckets<F><<<(unsigned)(((size_t)NWIN * NBUCKET + 127) / 128), 128, 0, st>>>((const u64 *)mont.p, n, (const u32 *)counts.p, (const u32 *)offsets.p, (const u32 *)sorted.p, (xyzz *)buckets.p); LAUNCHED(1); }
        { PhaseTimer t(ctx, "msm_windows");
          k_window_chunks<F><<<(NWIN * NCHUNK + 63) / 64, 64, 0, st>>>((const xyzz *)buckets.p, (xyzz *)chunks.p); LAUNCHED(1);
          k_window_final<F><<<NWIN, 128, 0, st>>>((const xyzz *)chunks.p, (xyzz *)wins.p); LAUNCHED(1); }
        std::vector<xyzz> w(NWIN);
        CUDA_TRY(cudaMemcpyAsync(w.data(), wins.p, NWIN * sizeof(xyzz), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        xyzz acc = pt_identity<F>();                   // Horner over the windows on the host: sum_w 2^(16 w) S_w
        for (int wi = NWIN - 1; wi >= 0; wi--) {
            for (int k = 0; k < WBITS; k++) acc = pt_double(acc);
            pt_add(acc, w[wi]);
        }
        if (pt_is_identity(acc)) { for (int i = 0; i < 2 * NL; i++) out_affine[i] = 0; return B200_OK; }
        F zi2 = f_inv_host(acc.zz), zi3 = f_inv_host(acc.zzz);
        f_store(f_mul(acc.x, zi2), out_affine); f_store(f_mul(acc.y, zi3), out_affine + NL);
        return B200_OK;
    } catch (const CudaErr &e) {
        char buf[256]; snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d", (int)e.e, cudaGetErrorString(e.e), e.file, e.line);
        cudaGetLastError(); return ctx_fail(ctx, e.e == cudaErrorMemoryAllocation ? B200_ERR_OOM : B200_ERR_CUDA, buf);
    } catch (const ApiErr &e) { return ctx_fail(ctx, e.code, e.msg); }
}
template <class F>
static int msm_host(b200_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, uint64_t *out_affine, const char *who) {
    if (!ctx || !points || !scalars || !n) return B200_ERR_BAD_ARG;
    try {
        cudaSetDevice(ctx->device);
        const size_t pb = n * 16 * limbs_of<F>::N;
        DevBuf p(pb), s(n * 32);
        CUDA_TRY(cudaMemcpyAsync(p.p, points, pb, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(s.p, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
        return msm_dev<F>(ctx, p.u(), s.u(), n, out_affine, who);
    } catch (const CudaErr &e) { return ctx_fail(ctx, B200_ERR_CUDA, std::string("CUDA error in ") + who); }
}

}  // namespace

extern "C" int b200_msm_g1_dev(b200_ctx *ctx, const uint64_t *d_points, const uint64_t *d_scalars, size_t n, uint64_t out_affine[8]) {
    return msm_dev<fq>(ctx, d_points, d_scalars, n, out_affine, "b200_msm_g1");
}
extern "C" int b200_msm_g1(b200_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, uint64_t out_affine[8]) {
    return msm_host<fq>(ctx, points, scalars, n, out_affine, "b200_msm_g1");
}
extern "C" int b200_msm_g2_dev(b200_ctx *ctx, const uint64_t *d_points, const uint64_t *d_scalars, size_t n, uint64_t out_affine[16]) {
    return msm_dev<fq2>(ctx, d_points, d_scalars, n, out_affine, "b200_msm_g2");
}
extern "C" int b200_msm_g2(b200_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, uint64_t out_affine[16]) {
    return msm_host<fq2>(ctx, points, scalars, n, out_affine, "b200_msm_g2");
}
