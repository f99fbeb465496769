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
#include <memory>
#include <string.h>
#include <algorithm>

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
constexpr u32 HOT_BUCKET = 2048, HOT_SLICE = 256;   // buckets above HOT_BUCKET points are summed in slices of HOT_SLICE by separate threads

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
    if (cnt > HOT_BUCKET) return;                               // split over many threads by k_hot_partial / k_hot_reduce
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
// ---- hot buckets --------------------------------------------------------------------------------------------------
// Groth16 witness scalars are heavily skewed (most wires are 0, 1 or small): with one thread per bucket the bucket of digit 1
// in window 0 would own a constant fraction of all N points.  Buckets with more than HOT_BUCKET points are therefore cut into
// slices of HOT_SLICE points, every slice is summed by its own thread (k_hot_partial), and one warp per hot bucket adds the
// slice sums (k_hot_reduce).  The task list is built on the device (k_hot_find): no host round trip.
struct HotTask { u32 bucket, slice; };
__global__ void k_hot_find(const u32 *__restrict__ counts, u32 *__restrict__ ctr /* [0] = #hot, [1] = #tasks */, u32 *__restrict__ hot_bucket,
                           u32 *__restrict__ hot_tbase, HotTask *__restrict__ tasks, u32 max_hot, u32 max_tasks) {
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= NWIN * NBUCKET) return;
    const u32 cnt = counts[b];
    if (cnt <= HOT_BUCKET) return;
    const u32 nt = (cnt + HOT_SLICE - 1) / HOT_SLICE;
    const u32 slot = atomicAdd(&ctr[0], 1u), tb = atomicAdd(&ctr[1], nt);
    if (slot >= max_hot || tb + nt > max_tasks) return;          // cannot happen: the bounds are the pigeonhole limits
    hot_bucket[slot] = b; hot_tbase[slot] = tb;
    for (u32 j = 0; j < nt; j++) { tasks[tb + j].bucket = b; tasks[tb + j].slice = j; }
}
template <class F>
__global__ void __launch_bounds__(128) k_hot_partial(const u64 *__restrict__ points, size_t n, const u32 *__restrict__ counts, const u32 *__restrict__ offsets,
                                                     const u32 *__restrict__ sorted, const u32 *__restrict__ ctr, const HotTask *__restrict__ tasks,
                                                     xyzz_t<F> *__restrict__ partial) {
    constexpr int NL = limbs_of<F>::N;
    const u32 ntask = ctr[1];
    for (u32 t = blockIdx.x * blockDim.x + threadIdx.x; t < ntask; t += gridDim.x * blockDim.x) {
        const u32 b = tasks[t].bucket, sl = tasks[t].slice;
        const size_t win = b / NBUCKET;
        const u32 cnt = counts[b], off = offsets[b];
        const u32 k0 = sl * HOT_SLICE, k1 = min(cnt, k0 + HOT_SLICE);
        const u32 *idx = sorted + win * n + off;
        xyzz_t<F> acc = pt_identity<F>();
        for (u32 k = k0; k < k1; k++) {
            const u64 *p = points + 2 * NL * (size_t)idx[k];
            u64 l[2 * NL], any = 0;
#pragma unroll
            for (int i = 0; i < 2 * NL; i++) { l[i] = p[i]; any |= l[i]; }
            if (any == 0) continue;
            F x, y; f_load_mont(x, l); f_load_mont(y, l + NL);
            pt_madd(acc, x, y);
        }
        partial[t] = acc;
    }
}
template <class F>
__global__ void __launch_bounds__(128) k_hot_reduce(const u32 *__restrict__ counts, const u32 *__restrict__ ctr, const u32 *__restrict__ hot_bucket,
                                                    const u32 *__restrict__ hot_tbase, const xyzz_t<F> *__restrict__ partial, xyzz_t<F> *__restrict__ buckets) {
    typedef xyzz_t<F> xyzz;
    __shared__ xyzz sh[128];
    const u32 nhot = ctr[0];
    for (u32 h = blockIdx.x; h < nhot; h += gridDim.x) {       // one block per hot bucket
        const u32 b = hot_bucket[h], tb = hot_tbase[h], nt = (counts[b] + HOT_SLICE - 1) / HOT_SLICE;
        xyzz s = pt_identity<F>();
        for (u32 j = threadIdx.x; j < nt; j += 128) pt_add(s, partial[tb + j]);
        sh[threadIdx.x] = s;
        __syncthreads();
        for (int off = 64; off > 0; off >>= 1) {
            if ((int)threadIdx.x < off) { xyzz a = sh[threadIdx.x]; pt_add(a, sh[threadIdx.x + off]); sh[threadIdx.x] = a; }
            __syncthreads();
        }
        if (threadIdx.x == 0) buckets[b] = sh[0];
        __syncthreads();
    }
}
// Horner over the windows on the device: sum_w 2^(16 w) S_w  (one thread: 256 doublings + 16 additions)
template <class F>
__global__ void k_horner(const xyzz_t<F> *__restrict__ wins, xyzz_t<F> *__restrict__ out) {
    if (blockIdx.x || threadIdx.x) return;
    xyzz_t<F> acc = pt_identity<F>();
    for (int wi = NWIN - 1; wi >= 0; wi--) {
        for (int k = 0; k < WBITS; k++) acc = pt_double(acc);
        pt_add(acc, wins[wi]);
    }
    out[0] = acc;
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
        // hot-bucket scratch: at most N*NWIN/HOT_BUCKET hot buckets and N*NWIN/HOT_SLICE + that many slices (pigeonhole)
        const u32 max_hot = (u32)std::min<size_t>((size_t)NWIN * NBUCKET, n * NWIN / HOT_BUCKET + 1), max_tasks = (u32)(n * NWIN / HOT_SLICE + max_hot + 1);
        const size_t sz_hot = rnd((size_t)max_hot * 8 + 64), sz_tasks = rnd((size_t)max_tasks * sizeof(HotTask)), sz_partial = rnd((size_t)max_tasks * sizeof(xyzz));
        u8 *base = ctx_arena(ctx, sz_digits + 3 * sz_cnt + sz_sorted + sz_buckets + sz_chunks + 2 * sz_wins + sz_mont + sz_hot + sz_tasks + sz_partial);
        struct Buf { void *p; size_t bytes; };
        size_t off = 0; auto take = [&](size_t b) { Buf r{base + off, b}; off += b; return r; };
        Buf digits = take(sz_digits), counts = take(sz_cnt), offsets = take(sz_cnt), cursor = take(sz_cnt), sorted = take(sz_sorted),
            buckets = take(sz_buckets), chunks = take(sz_chunks), wins = take(sz_wins), total = take(sz_wins), mont = take(sz_mont),
            hot = take(sz_hot), tasks = take(sz_tasks), partial = take(sz_partial);
        u32 *d_ctr = (u32 *)hot.p, *d_hot_bucket = d_ctr + 16, *d_hot_tbase = d_hot_bucket + max_hot;
        CUDA_TRY(cudaMemsetAsync(d_ctr, 0, 64, st));
        CUDA_TRY(cudaMemsetAsync(counts.p, 0, counts.bytes, st)); CUDA_TRY(cudaMemsetAsync(cursor.p, 0, cursor.bytes, st));
        { PhaseTimer t(ctx, "msm_to_mont"); k_points_to_mont<<<(unsigned)((n * (NL / 2) + 255) / 256), 256, 0, st>>>(d_points, (u64 *)mont.p, n * (NL / 2)); LAUNCHED(1); }
        { PhaseTimer t(ctx, "msm_digits"); k_digits<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_scalars, n, (u32 *)digits.p, (u32 *)counts.p); LAUNCHED(1); }
        { PhaseTimer t(ctx, "msm_sort");
          k_scan<<<NWIN, 1024, 0, st>>>((const u32 *)counts.p, (u32 *)offsets.p); LAUNCHED(1);
          k_scatter<<<(unsigned)((n * NWIN + 255) / 256), 256, 0, st>>>((const u32 *)digits.p, n, (const u32 *)offsets.p, (u32 *)cursor.p, (u32 *)sorted.p); LAUNCHED(1); }
        { PhaseTimer t(ctx, "msm_buckets");
          k_hot_find<<<(NWIN * NBUCKET + 255) / 256, 256, 0, st>>>((const u32 *)counts.p, d_ctr, d_hot_bucket, d_hot_tbase, (HotTask *)tasks.p, max_hot, max_tasks); LAUNCHED(1);
          k_buckets<F><<<(unsigned)(((size_t)NWIN * NBUCKET + 127) / 128), 128, 0, st>>>((const u64 *)mont.p, n, (const u32 *)counts.p, (const u32 *)offsets.p, (const u32 *)sorted.p, (xyzz *)buckets.p); LAUNCHED(1);
          k_hot_partial<F><<<(unsigned)std::min<size_t>(((size_t)max_tasks + 127) / 128, 4096), 128, 0, st>>>((const u64 *)mont.p, n, (const u32 *)counts.p, (const u32 *)offsets.p, (const u32 *)sorted.p, d_ctr, (const HotTask *)tasks.p, (xyzz *)partial.p); LAUNCHED(1);
          k_hot_reduce<F><<<256, 128, 0, st>>>((const u32 *)counts.p, d_ctr, d_hot_bucket, d_hot_tbase, (const xyzz *)partial.p, (xyzz *)buckets.p); LAUNCHED(1); }
        { PhaseTimer t(ctx, "msm_windows");
          k_window_chunks<F><<<(NWIN * NCHUNK + 63) / 64, 64, 0, st>>>((const xyzz *)buckets.p, (xyzz *)chunks.p); LAUNCHED(1);
          k_window_final<F><<<NWIN, 128, 0, st>>>((const xyzz *)chunks.p, (xyzz *)wins.p); LAUNCHED(1); }
        k_horner<F><<<1, 32, 0, st>>>((const xyzz *)wins.p, (xyzz *)total.p); LAUNCHED(1);
        xyzz acc;
        CUDA_TRY(cudaMemcpyAsync(&acc, total.p, sizeof(xyzz), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
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


// =====================================================================================================================
// groth16.Prove as a pipeline (gnark-plonky2-verifier/cmd/web-api.go:77 -> gnark v0.9.1
// backend/groth16/bn254/prove.go, not vendored).  Restated from the published Groth16 prover with gnark's conventions:
//   H = (A.B - C) / Z on the coset 5<w>:  per vector an inverse FFT and a coset FFT, the pointwise quotient (Z is the
//       constant 5^n - 1 there), a coset inverse FFT                                          -> b200_fr_ntt_dev (fr_ntt.cu) x 7
//   Ar  = alpha + sum_i w_i A_i + r delta          Bs = beta + sum_i w_i B_i + s delta  (in G1 and in G2)
//   Krs = sum_{i private} w_i K_i + sum_k h_k Z_k + s Ar + r Bs1 - r s delta                   -> b200_msm_g1/g2 (above) x 5
//   proof = (Ar, Bs in G2, Krs).  r, s are the caller's blinding scalars (gnark draws them at random, so proof BYTES are
//   not reproducible across runs by design; tests check the discrete logarithms against a toxic-waste key).
// The proving key is uploaded once (b200_groth16_pk_upload) and stays device-resident.
// =====================================================================================================================
#include "bn254.cuh"
int fr_ntt_dev_internal(b200_ctx *ctx, const uint64_t *d_in, uint32_t log_n, int inverse, int coset, uint64_t *d_out);   // fr_ntt.cu

struct b200_groth16_pk {
    b200_ctx *ctx; u32 log_domain; size_t num_wires, num_public;
    DevBuf g1_a, g1_b, g2_b, g1_k, g1_z;
    u64 g1_alpha[8], g1_beta[8], g1_delta[8], g2_beta[16], g2_delta[16];
};

namespace {
// H[i] = (A[i] * B[i] - C[i]) * zinv, canonical limbs in and out (zinv_m: Montgomery form)
__global__ void __launch_bounds__(128) k_groth16_h(const u64 *__restrict__ a, const u64 *__restrict__ b, const u64 *__restrict__ c, u64 *__restrict__ h, size_t n, fr_t zinv_m) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 la[4], lb[4], lc[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { la[k] = a[4 * i + k]; lb[k] = b[4 * i + k]; lc[k] = c[4 * i + k]; }
    fr_t x = fr_sub(fr_mul(fr_from_canonical(la), fr_from_canonical(lb)), fr_from_canonical(lc));
    x = fr_mul(x, zinv_m);
    u64 o[4]; fr_to_canonical(x, o);
#pragma unroll
    for (int k = 0; k < 4; k++) h[4 * i + k] = o[k];
}
static fr_t fr_inv_host(const fr_t &a) {          // Fermat
    u64 e[4] = {FR_M0 - 2, FR_M1, FR_M2, FR_M3};
    const u64 one_c[4] = {1, 0, 0, 0};
    fr_t r = fr_from_canonical(one_c), b = a;
    for (int i = 0; i < 256; i++) { if ((e[i / 64] >> (i % 64)) & 1) r = fr_mul(r, b); b = fr_mul(b, b); }
    return r;
}
static fr_t fr_pow_host(fr_t b, u64 e) { const u64 one_c[4] = {1, 0, 0, 0}; fr_t r = fr_from_canonical(one_c); while (e) { if (e & 1) r = fr_mul(r, b); b = fr_mul(b, b); e >>= 1; } return r; }

// host group arithmetic on affine canonical limbs (final combination only: a handful of operations)
template <class F> static xyzz_t<F> pt_from_affine(const u64 *p) {
    constexpr int NL = limbs_of<F>::N;
    u64 any = 0; for (int i = 0; i < 2 * NL; i++) any |= p[i];
    if (!any) return pt_identity<F>();
    xyzz_t<F> r;
    if constexpr (NL == 4) { r.x = fq_to_mont(p); r.y = fq_to_mont(p + 4); }
    else { r.x.a = fq_to_mont(p); r.x.b = fq_to_mont(p + 4); r.y.a = fq_to_mont(p + 8); r.y.b = fq_to_mont(p + 12); }
    f_set_one(r.zz); f_set_one(r.zzz);
    return r;
}
template <class F> static void pt_to_affine(const xyzz_t<F> &a, u64 *out) {
    constexpr int NL = limbs_of<F>::N;
    if (pt_is_identity(a)) { for (int i = 0; i < 2 * NL; i++) out[i] = 0; return; }
    F zi2 = f_inv_host(a.zz), zi3 = f_inv_host(a.zzz);
    f_store(f_mul(a.x, zi2), out); f_store(f_mul(a.y, zi3), out + NL);
}
template <class F> static xyzz_t<F> pt_scalar_mul_host(const xyzz_t<F> &p, const u64 k[4]) {
    xyzz_t<F> acc = pt_identity<F>();
    for (int i = 255; i >= 0; i--) { acc = pt_double(acc); if ((k[i / 64] >> (i % 64)) & 1) pt_add(acc, p); }
    return acc;
}
template <class F> static xyzz_t<F> pt_neg(xyzz_t<F> p) { F z; f_set_zero(z); p.y = f_sub(z, p.y); return p; }
}  // namespace

extern "C" int b200_groth16_pk_upload(b200_ctx *ctx, const b200_groth16_pk_desc *d, b200_groth16_pk **out) {
    if (!ctx) return B200_ERR_BAD_ARG;
    try {
        cudaSetDevice(ctx->device);
        if (!d || !out || !d->g1_a || !d->g1_b || !d->g2_b || !d->g1_z || d->log_domain == 0 || d->log_domain > 28 || d->num_public > d->num_wires || d->num_wires == 0 ||
            (d->num_wires > d->num_public && !d->g1_k))
            throw ApiErr{B200_ERR_BAD_ARG, "b200_groth16_pk_upload: bad arguments"};
        *out = nullptr;
        std::unique_ptr<b200_groth16_pk> pk(new b200_groth16_pk());
        pk->ctx = ctx; pk->log_domain = d->log_domain; pk->num_wires = d->num_wires; pk->num_public = d->num_public;
        const size_t nw = d->num_wires, npriv = nw - d->num_public, nz = (size_t(1) << d->log_domain) - 1;
        auto up = [&](DevBuf &b, const u64 *src, size_t bytes) { b.alloc(std::max<size_t>(bytes, 8)); if (bytes) CUDA_TRY(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream)); };
        up(pk->g1_a, d->g1_a, nw * 64); up(pk->g1_b, d->g1_b, nw * 64); up(pk->g2_b, d->g2_b, nw * 128); up(pk->g1_k, d->g1_k, npriv * 64); up(pk->g1_z, d->g1_z, nz * 64);
        memcpy(pk->g1_alpha, d->g1_alpha, 64); memcpy(pk->g1_beta, d->g1_beta, 64); memcpy(pk->g1_delta, d->g1_delta, 64);
        memcpy(pk->g2_beta, d->g2_beta, 128); memcpy(pk->g2_delta, d->g2_delta, 128);
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        *out = pk.release();
        return B200_OK;
    } catch (const CudaErr &e) { return ctx_fail(ctx, e.e == cudaErrorMemoryAllocation ? B200_ERR_OOM : B200_ERR_CUDA, "CUDA error in b200_groth16_pk_upload");
    } catch (const ApiErr &e) { return ctx_fail(ctx, e.code, e.msg); }
}
extern "C" void b200_groth16_pk_free(b200_groth16_pk *pk) { if (!pk) return; cudaSetDevice(pk->ctx->device); cudaStreamSynchronize(pk->ctx->stream); delete pk; }

extern "C" int b200_groth16_prove(b200_ctx *ctx, const b200_groth16_pk *pk, const uint64_t *wires, const uint64_t *a, const uint64_t *b, const uint64_t *c,
                                  const uint64_t r[4], const uint64_t s[4], uint64_t proof[32]) {
    if (!ctx) return B200_ERR_BAD_ARG;
    try {
        cudaSetDevice(ctx->device);
        if (!pk || pk->ctx != ctx || !wires || !a || !b || !c || !r || !s || !proof) throw ApiErr{B200_ERR_BAD_ARG, "b200_groth16_prove: bad arguments"};
        cudaStream_t st = ctx->stream;
        const size_t n = size_t(1) << pk->log_domain, nw = pk->num_wires, npriv = nw - pk->num_public;
        DevBuf dw(nw * 32), da(n * 32), db(n * 32), dc(n * 32), t0(n * 32), t1(n * 32);
        CUDA_TRY(cudaMemcpyAsync(dw.p, wires, nw * 32, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(da.p, a, n * 32, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(db.p, b, n * 32, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(dc.p, c, n * 32, cudaMemcpyHostToDevice, st));
        // ---- H = (A.B - C)/Z on the coset ----
        {
            PhaseTimer tm(ctx, "groth16_h");
            DevBuf *vec[3] = {&da, &db, &dc};
            for (DevBuf *v : vec) {                              // Lagrange values -> coefficients -> values on the coset 5<w>
                int rc = fr_ntt_dev_internal(ctx, v->u(), pk->log_domain, 1, 0, t0.u()); if (rc) return rc;
                rc = fr_ntt_dev_internal(ctx, t0.u(), pk->log_domain, 0, 1, v->u()); if (rc) return rc;
            }
            const u64 five[4] = {5, 0, 0, 0}, one_c[4] = {1, 0, 0, 0};
            fr_t zc = fr_sub(fr_pow_host(fr_from_canonical(five), (u64)n), fr_from_canonical(one_c));      // Z on the coset: 5^n - 1
            fr_t zinv = fr_inv_host(zc);
            k_groth16_h<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(da.u(), db.u(), dc.u(), t1.u(), n, zinv); LAUNCHED(1);
            int rc = fr_ntt_dev_internal(ctx, t1.u(), pk->log_domain, 1, 1, t0.u()); if (rc) return rc;        // h coefficients (natural order)
        }
        // ---- the five MSMs ----
        u64 Ar[8], Bs1[8], Bs2[16], Kr[8], Hz[8];
        int rc;
        if ((rc = msm_dev<fq>(ctx, pk->g1_a.u(), dw.u(), nw, Ar, "b200_groth16_prove (A)"))) return rc;
        if ((rc = msm_dev<fq>(ctx, pk->g1_b.u(), dw.u(), nw, Bs1, "b200_groth16_prove (B1)"))) return rc;
        if ((rc = msm_dev<fq2>(ctx, pk->g2_b.u(), dw.u(), nw, Bs2, "b200_groth16_prove (B2)"))) return rc;
        if (npriv) { if ((rc = msm_dev<fq>(ctx, pk->g1_k.u(), dw.u() + 4 * pk->num_public, npriv, Kr, "b200_groth16_prove (K)"))) return rc; }
        else memset(Kr, 0, sizeof Kr);
        if ((rc = msm_dev<fq>(ctx, pk->g1_z.u(), t0.u(), n - 1, Hz, "b200_groth16_prove (Z)"))) return rc;
        // ---- blinding and assembly (host: ~1 k field products) ----
        typedef xyzz_t<fq> P1; typedef xyzz_t<fq2> P2;
        P1 delta1 = pt_from_affine<fq>(pk->g1_delta), ar = pt_from_affine<fq>(pk->g1_alpha), bs1 = pt_from_affine<fq>(pk->g1_beta);
        pt_add(ar, pt_from_affine<fq>(Ar)); pt_add(ar, pt_scalar_mul_host(delta1, r));
        pt_add(bs1, pt_from_affine<fq>(Bs1)); pt_add(bs1, pt_scalar_mul_host(delta1, s));
        P2 delta2 = pt_from_affine<fq2>(pk->g2_delta), bs2 = pt_from_affine<fq2>(pk->g2_beta);
        pt_add(bs2, pt_from_affine<fq2>(Bs2)); pt_add(bs2, pt_scalar_mul_host(delta2, s));
        u64 rs[4]; { fr_t m = fr_mul(fr_from_canonical(r), fr_from_canonical(s)); fr_to_canonical(m, rs); }
        P1 krs = pt_from_affine<fq>(Kr);
        pt_add(krs, pt_from_affine<fq>(Hz)); pt_add(krs, pt_scalar_mul_host(ar, s)); pt_add(krs, pt_scalar_mul_host(bs1, r));
        pt_add(krs, pt_neg(pt_scalar_mul_host(delta1, rs)));
        pt_to_affine(ar, proof); pt_to_affine(bs2, proof + 8); pt_to_affine(krs, proof + 24);
        return B200_OK;
    } catch (const CudaErr &e) {
        char buf[256]; snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d", (int)e.e, cudaGetErrorString(e.e), e.file, e.line);
        cudaGetLastError(); return ctx_fail(ctx, e.e == cudaErrorMemoryAllocation ? B200_ERR_OOM : B200_ERR_CUDA, buf);
    } catch (const ApiErr &e) { return ctx_fail(ctx, e.code, e.msg); }
}
