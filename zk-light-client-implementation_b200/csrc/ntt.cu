// Batched Goldilocks radix-2 DIF NTT for sm_100a: shared-memory tiles, radix-8 butterflies in registers.
// Replaces upstream plonky2_field::fft::{fft_dispatch, ifft} and PolynomialValues::lde / coset_fft as used by
// PolynomialBatch::from_values / from_coeffs (loop A of section 8).
//
// Data layout: column-major batches, one polynomial = 2^L contiguous u64.  A transform of 2^L points is split
// into passes of B <= 12 layers.  A pass works on tiles of 2^B x C elements staged in shared memory (36 KB with
// padding): the "strided" passes (span S >= 16) take C = 16 consecutive low indices so that every global access is a
// full 128-byte line; the last pass (S = 1) takes 2^B contiguous points.  Inside a pass each thread pulls 8 elements
// into registers and performs three butterfly layers (12 modmuls) per shared-memory round trip.  Twiddles come from
// one table  tw[h + j] = w_{2h}^j  shared by all sizes (L2-resident).
#include "b200_internal.h"

#define NTT_THREADS 256
#define NTT_MAX_TILE_LOG 12
#define SM_IDX(i) ((i) + ((i) >> 3))       // one pad word per 8 elements: conflict-free radix-8 access

__global__ void k_gen_twiddles(u64 *fwd, u64 *inv, int max_log) {
    size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t n = size_t(1) << max_log;
    if (idx >= n || idx == 0) return;
    int s = 63 - __clzll((long long)idx);            // idx = h + j, h = 2^s
    u64 h = u64(1) << s, j = idx - h;
    u64 w = gl_root_of_unity(s + 1);
    fwd[idx] = gl_pow(w, j);
    inv[idx] = gl_pow(gl_inv(w), j);
}

void ntt_tables_ensure(NttTables &T, int log_n, cudaStream_t st) {
    if (log_n <= T.max_log) return;
    if (log_n < 16) log_n = 16;
    ntt_tables_free(T);
    size_t n = size_t(1) << log_n;
    CUDA_TRY(cudaMalloc(&T.fwd, n * sizeof(u64)));
    CUDA_TRY(cudaMalloc(&T.inv, n * sizeof(u64)));
    k_gen_twiddles<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(T.fwd, T.inv, log_n);
    LAUNCHED(1);
    T.max_log = log_n;
}
void ntt_tables_free(NttTables &T) {
    if (T.fwd) cudaFree(T.fwd);
    if (T.inv) cudaFree(T.inv);
    T.fwd = T.inv = nullptr; T.max_log = 0;
}

struct PassArgs {
    const u64 *in; u64 *out; const u64 *tw; const u64 *scale;
    size_t in_stride, out_stride;
    int L, l_begin, B, logC, rate_bits, in_bitrev;
    u64 out_mul;
    unsigned tiles_per_item;
};

// R-point DIF butterflies in registers.  x[e] sits at mid index (m_hi << (q+r)) | (e << q) | m_lo.
template <int r>
__device__ __forceinline__ void radix_dif(u64 *x, const u64 *tw, u64 span_unit, u64 jbase, int q) {
    // span_unit = S (global stride of one mid step), jbase = m_lo * S + low
#pragma unroll
    for (int s = r - 1; s >= 0; s--) {
        u64 h = (u64(1) << (q + s)) * span_unit;
#pragma unroll
        for (int e = 0; e < (1 << r); e++) {
            if (e & (1 << s)) continue;
            u64 j = jbase + (u64)(e & ((1 << s) - 1)) * (span_unit << q);      // (i mod h) of the upper element
            u64 w = __ldg(tw + h + j);
            u64 u = x[e], v = x[e + (1 << s)];
            x[e] = gl_add(u, v);
            x[e + (1 << s)] = gl_mul(gl_sub(u, v), w);
        }
    }
}

template <int r>
__device__ __forceinline__ void do_round(u64 *sm, const PassArgs &a, int q, u64 S, u64 low0, bool strided) {
    const int B = a.B, C = 1 << a.logC;
    const int tasks = (1 << (B - r)) * C;
    for (int t = threadIdx.x; t < tasks; t += NTT_THREADS) {
        int c, m_lo, m_hi;
        if (strided) { c = t & (C - 1); int rest = t >> a.logC; m_lo = rest & ((1 << q) - 1); m_hi = rest >> q; }
        else { m_lo = t & ((1 << q) - 1); int rest = t >> q; m_hi = rest & ((1 << (B - r - q)) - 1); c = rest >> (B - r - q); }
        u64 x[1 << r];
        int base_mid = (m_hi << (q + r)) | m_lo;
#pragma unroll
        for (int e = 0; e < (1 << r); e++) {
            int mid = base_mid | (e << q);
            int i = strided ? (mid << a.logC) + c : (c << B) + mid;
            x[e] = sm[SM_IDX(i)];
        }
        u64 jbase = (u64)m_lo * S + (strided ? low0 + c : 0);
        radix_dif<r>(x, a.tw, S, jbase, q);
#pragma unroll
        for (int e = 0; e < (1 << r); e++) {
            int mid = base_mid | (e << q);
            int i = strided ? (mid << a.logC) + c : (c << B) + mid;
            sm[SM_IDX(i)] = x[e];
        }
    }
}

__global__ void __launch_bounds__(NTT_THREADS) k_ntt_dif_pass(PassArgs a) {
    extern __shared__ u64 sm[];
    const int B = a.B, C = 1 << a.logC, L = a.L;
    const int T = (1 << B) * C;
    const size_t item = blockIdx.x / a.tiles_per_item;
    const unsigned tile = blockIdx.x % a.tiles_per_item;
    const size_t col = item >> a.rate_bits;
    const unsigned rr = (unsigned)(item & ((size_t(1) << a.rate_bits) - 1));
    const u64 N = u64(1) << L;
    const u64 *in = a.in + col * a.in_stride;
    u64 *out = a.out + col * a.out_stride + (u64)bitrev32(rr, a.rate_bits) * N;
    const u64 *scale = a.scale ? a.scale + (u64)rr * N : nullptr;
    const int lend = a.l_begin + B;
    const u64 S = u64(1) << (L - lend);                 // global stride of one `mid` step
    const bool strided = S > 1;
    // tile -> (top, low0): strided: tiles enumerate (top, low-group); contiguous: tiles enumerate groups of C tops
    u64 top, low0;
    if (strided) { u64 lows = S >> a.logC; top = tile / lows; low0 = (tile % lows) << a.logC; }
    else { top = (u64)tile << a.logC; low0 = 0; }
    const u64 top_base = top << (L - a.l_begin);        // strided: index of (top, mid=0, low=0); contiguous: first group base

    // ---- load ----
    if (!a.in_bitrev) {
        for (int idx = threadIdx.x; idx < T; idx += NTT_THREADS) {
            u64 g; int i;
            if (strided) { int c = idx & (C - 1), mid = idx >> a.logC; g = top_base + (u64)mid * S + low0 + c; i = idx; }
            else { g = top_base + idx; i = idx; }
            u64 v = in[g];
            if (scale) v = gl_mul(v, scale[g]);
            sm[SM_IDX(i)] = v;
        }
    } else {
        // l_begin == 0 (top == 0).  natural index i = mid*S + low lives at storage position bitrev_L(i).
        for (int idx = threadIdx.x; idx < T; idx += NTT_THREADS) {
            int midr = idx & ((1 << B) - 1), c = idx >> B;
            u64 pos; int i;
            if (strided) { u64 low = low0 + c; pos = ((u64)bitrev32((u32)low, L - B) << B) + midr; i = ((int)bitrev32(midr, B) << a.logC) + c; }
            else { pos = ((u64)bitrev32((u32)(top + c), L - B) << B) + midr; i = (c << B) + (int)bitrev32(midr, B); }
            u64 v = in[pos];
            if (scale) v = gl_mul(v, scale[pos]);
            sm[SM_IDX(i)] = v;
        }
    }
    __syncthreads();
    // ---- butterfly rounds, high mid bits first ----
    int q = B;
    const int rem = B % 3;
    if (rem == 1) { q -= 1; do_round<1>(sm, a, q, S, low0, strided); __syncthreads(); }
    else if (rem == 2) { q -= 2; do_round<2>(sm, a, q, S, low0, strided); __syncthreads(); }
    while (q > 0) { q -= 3; do_round<3>(sm, a, q, S, low0, strided); __syncthreads(); }
    // ---- store ----
    const bool mul = a.out_mul != 1;
    for (int idx = threadIdx.x; idx < T; idx += NTT_THREADS) {
        u64 g;
        if (strided) { int c = idx & (C - 1), mid = idx >> a.logC; g = top_base + (u64)mid * S + low0 + c; }
        else g = top_base + idx;
        u64 v = sm[SM_IDX(idx)];
        if (mul) v = gl_mul(v, a.out_mul);
        out[g] = v;
    }
}

void ntt_dif(const NttTables &T, const u64 *in, size_t in_stride, bool in_bitrev, u64 *out, size_t out_stride,
             int L, size_t items, int rate_bits, const u64 *scale, bool inverse, u64 out_mul, cudaStream_t st) {
    if (L > T.max_log) throw CudaErr{cudaErrorInvalidValue, __FILE__, __LINE__};
    if (L == 0) {   // size-1 transform: copy
        throw CudaErr{cudaErrorInvalidValue, __FILE__, __LINE__};
    }
    // 2^16 .. 2^20 points: bulk-copy staged passes with register-resident twiddles (ntt2.cu)
    if (ntt2_supported(L, in_stride, out_stride, rate_bits) && (in_bitrev || (rate_bits == 0 && scale == nullptr)) &&
        (!rate_bits || out_stride == (size_t(1) << (L + rate_bits))) && ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
        (!scale || (uintptr_t)scale % 16 == 0)) {
        ntt2_dif(T, in, in_stride, in_bitrev, out, out_stride, L, items, rate_bits, scale, inverse, out_mul, st);
        return;
    }
    // split L layers into passes: all strided passes must have S >= 16
    int npass = (L + NTT_MAX_TILE_LOG - 1) / NTT_MAX_TILE_LOG;
    std::vector<int> Bs(npass);
    {   // last pass gets the largest share (<= 12) ; earlier passes use B = 8 (tile 2^8 x 16)
        int left = L;
        for (int p = npass - 1; p >= 0; p--) {
            int b = (left + p) / (p + 1);           // balanced
            if (p < npass - 1 && b > 8) b = 8;      // strided passes: 2^8 x 16 tile
            Bs[p] = b; left -= b;
        }
        // rebalance: if the earlier passes were clamped the last one absorbs the rest
        int sum = 0; for (int b : Bs) sum += b;
        Bs[npass - 1] += L - sum;
        while (Bs[npass - 1] > NTT_MAX_TILE_LOG) {  // need one more pass
            Bs.insert(Bs.begin(), 8); Bs[Bs.size() - 1] -= 8; npass++;
        }
    }
    int l_begin = 0;
    for (int p = 0; p < npass; p++) {
        PassArgs a;
        bool first = (p == 0), last = (p == npass - 1);
        a.in = first ? in : out; a.out = out; a.tw = inverse ? T.inv : T.fwd;
        a.scale = first ? scale : nullptr;
        a.in_stride = first ? in_stride : out_stride; a.out_stride = out_stride;
        a.L = L; a.l_begin = l_begin; a.B = Bs[p]; a.rate_bits = first ? rate_bits : 0;
        a.in_bitrev = first && in_bitrev;
        a.out_mul = last ? out_mul : 1;
        int lend = l_begin + Bs[p];
        bool strided = (L - lend) > 0;
        if (strided) a.logC = 4;
        else { a.logC = NTT_MAX_TILE_LOG - Bs[p]; if (a.logC > L - Bs[p]) a.logC = L - Bs[p]; if (a.logC > 4) a.logC = 4; }
        size_t tile_elems = size_t(1) << (a.B + a.logC);
        a.tiles_per_item = (unsigned)((size_t(1) << L) / tile_elems);
        size_t n_items = first ? items : (rate_bits ? items : items);
        // after the first pass of an LDE the 2^rate_bits coset blocks of a column are contiguous in `out`
        if (!first && rate_bits) { a.out_stride = size_t(1) << L; a.in_stride = a.out_stride; }
        if (!first && rate_bits && out_stride != (size_t(1) << (L + rate_bits))) throw CudaErr{cudaErrorInvalidValue, __FILE__, __LINE__};
        size_t blocks = n_items * a.tiles_per_item;
        size_t smem = (tile_elems + (tile_elems >> 3) + 8) * sizeof(u64);
        k_ntt_dif_pass<<<(unsigned)blocks, NTT_THREADS, smem, st>>>(a);
        LAUNCHED(1);
        l_begin = lend;
    }
}

__global__ void k_bitrev_permute(const u64 *in, u64 *out, int L, size_t total) {
    size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= total) return;
    size_t it = idx >> L; u32 i = (u32)(idx & ((size_t(1) << L) - 1));
    out[(it << L) + bitrev32(i, L)] = in[idx];
}
// tiled variant for L >= 10: index = hi(5) | mid | lo(5); the 32x32 tile (hi, lo) of a fixed mid lands at (rev lo, rev hi) of
// rev(mid): both the load (over lo) and the store (over rev hi) touch 256-byte runs.
__global__ void __launch_bounds__(256) k_bitrev_tiled(const u64 *__restrict__ in, u64 *__restrict__ out, int L, size_t tiles_per_item) {
    __shared__ u64 tile[32][33];
    const size_t it = blockIdx.x / tiles_per_item;
    const u32 mid = (u32)(blockIdx.x % tiles_per_item);
    const int mbits = L - 10;
    const u64 *src = in + (it << L); u64 *dst = out + (it << L);
    const u32 lane = threadIdx.x & 31, row0 = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        u32 hi = row0 + 8 * k;
        tile[hi][lane] = src[((size_t)hi << (L - 5)) | ((size_t)mid << 5) | lane];
    }
    __syncthreads();
    const u32 rmid = bitrev32(mid, mbits);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        u32 r = row0 + 8 * k;                                   // r = rev(lo), lane = rev(hi)
        dst[((size_t)r << (L - 5)) | ((size_t)rmid << 5) | lane] = tile[bitrev32(lane, 5)][bitrev32(r, 5)];
    }
}
void bitrev_permute(const u64 *in, u64 *out, int L, size_t items, cudaStream_t st) {
    if (L >= 10) {
        size_t tiles = size_t(1) << (L - 10);
        k_bitrev_tiled<<<(unsigned)(items * tiles), 256, 0, st>>>(in, out, L, tiles);
        LAUNCHED(1);
        return;
    }
    size_t total = items << L;
    k_bitrev_permute<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, out, L, total);
    LAUNCHED(1);
}

__global__ void k_lde_scale(u64 *out, int L, int rate_bits, u64 shift) {
    size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t total = size_t(1) << (L + rate_bits);
    if (idx >= total) return;
    u32 r = (u32)(idx >> L), p = (u32)(idx & ((size_t(1) << L) - 1));
    u64 base = gl_mul(shift, gl_pow(gl_root_of_unity(L + rate_bits), r));
    out[idx] = gl_pow(base, bitrev32(p, L));
}
void lde_scale_table(u64 *out, int L, int rate_bits, u64 shift, cudaStream_t st) {
    size_t total = size_t(1) << (L + rate_bits);
    k_lde_scale<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(out, L, rate_bits, shift);
    LAUNCHED(1);
}
