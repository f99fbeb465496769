// Poseidon Merkle-cap commitment kernels (loop B of SURVEY.md section 8 / a7).
// Replaces upstream plonky2::hash::merkle_tree::MerkleTree::new + PoseidonHash::{hash_or_noop, two_to_one}.
//
// Leaves are never materialised row-major: the LDE stays column-major [cols][L] (point index bit-reversed), so one
// thread per leaf walks the columns and every load instruction of a warp reads 32 consecutive u64 (256 B).  A leaf of
// `cols` elements costs ceil(cols/8) permutations (identity "hash" when cols <= 4); inner nodes cost one each.
// Digest storage: level 0 (L x 4 u64) | level 1 (L/2 x 4) | ... | cap level (2^cap_height x 4).
#include "b200_internal.h"
#include "poseidon.cuh"

#define MK_THREADS 128

// state of thread t lives in shared memory at sm[i * MK_THREADS + t], i < 12
#define ST(i) sm[(i) * MK_THREADS + threadIdx.x]
__global__ void __launch_bounds__(MK_THREADS) k_leaves_colmajor(const u64 *__restrict__ data, size_t cols, size_t L, u64 *__restrict__ digests) {
    __shared__ u64 sm[12 * MK_THREADS];
    size_t leaf = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const bool active = leaf < L;
    if (!active) leaf = L - 1;
#pragma unroll
    for (int i = 0; i < 12; i++) ST(i) = 0;
    if (cols <= 4) {
        for (size_t c = 0; c < cols; c++) ST(c) = data[c * L + leaf];
    } else {
        for (size_t c = 0; c < cols; c += 8) {
            u64 v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = (c + j < cols) ? __ldg(data + (c + j) * L + leaf) : 0;
#pragma unroll
            for (int j = 0; j < 8; j++) if (c + j < cols) ST(j) = v[j];
            poseidon_permute_smem(&ST(0), MK_THREADS);
        }
    }
    if (active) {
        ulonglong2 *o = reinterpret_cast<ulonglong2 *>(digests + 4 * leaf);
        o[0] = make_ulonglong2(ST(0), ST(1)); o[1] = make_ulonglong2(ST(2), ST(3));
    }
}

__global__ void __launch_bounds__(MK_THREADS) k_leaves_rowmajor(const u64 *__restrict__ data, size_t width, size_t L, u64 *__restrict__ digests) {
    __shared__ u64 sm[12 * MK_THREADS];
    size_t leaf = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const bool active = leaf < L;
    if (!active) leaf = L - 1;
    const u64 *row = data + leaf * width;
#pragma unroll
    for (int i = 0; i < 12; i++) ST(i) = 0;
    if (width <= 4) {
        for (size_t c = 0; c < width; c++) ST(c) = row[c];
    } else {
        for (size_t c = 0; c < width; c += 8) {
#pragma unroll
            for (int j = 0; j < 8; j++) if (c + j < width) ST(j) = row[c + j];
            poseidon_permute_smem(&ST(0), MK_THREADS);
        }
    }
    if (active) {
        ulonglong2 *o = reinterpret_cast<ulonglong2 *>(digests + 4 * leaf);
        o[0] = make_ulonglong2(ST(0), ST(1)); o[1] = make_ulonglong2(ST(2), ST(3));
    }
}

__global__ void __launch_bounds__(MK_THREADS) k_merkle_level(const u64 *__restrict__ in, u64 *__restrict__ out, size_t n_out) {
    __shared__ u64 sm[12 * MK_THREADS];
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const bool active = i < n_out;
    if (!active) i = n_out - 1;
    const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(in + 8 * i);
    ulonglong2 a = p[0], b = p[1], c = p[2], d = p[3];
    ST(0) = a.x; ST(1) = a.y; ST(2) = b.x; ST(3) = b.y; ST(4) = c.x; ST(5) = c.y; ST(6) = d.x; ST(7) = d.y;
    ST(8) = 0; ST(9) = 0; ST(10) = 0; ST(11) = 0;
    poseidon_permute_smem(&ST(0), MK_THREADS);
    if (active) {
        ulonglong2 *o = reinterpret_cast<ulonglong2 *>(out + 4 * i);
        o[0] = make_ulonglong2(ST(0), ST(1)); o[1] = make_ulonglong2(ST(2), ST(3));
    }
}

void merkle_leaves_colmajor(const u64 *data, size_t cols, int log_L, u64 *digests, cudaStream_t st) {
    size_t L = size_t(1) << log_L;
    k_leaves_colmajor<<<(unsigned)((L + MK_THREADS - 1) / MK_THREADS), MK_THREADS, 0, st>>>(data, cols, L, digests);
    LAUNCHED(1);
}
void merkle_leaves_rowmajor(const u64 *data, size_t width, int log_L, u64 *digests, cudaStream_t st) {
    size_t L = size_t(1) << log_L;
    k_leaves_rowmajor<<<(unsigned)((L + MK_THREADS - 1) / MK_THREADS), MK_THREADS, 0, st>>>(data, width, L, digests);
    LAUNCHED(1);
}
void merkle_build(u64 *digests, int log_L, int cap_height, cudaStream_t st) {
    u64 *cur = digests;
    for (int lv = log_L; lv > cap_height; lv--) {
        size_t n = size_t(1) << lv;
        u64 *nxt = cur + 4 * n;
        k_merkle_level<<<(unsigned)((n / 2 + MK_THREADS - 1) / MK_THREADS), MK_THREADS, 0, st>>>(cur, nxt, n / 2);
        LAUNCHED(1);
        cur = nxt;
    }
}
