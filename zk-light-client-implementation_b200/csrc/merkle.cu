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

__global__ void __launch_bounds__(MK_THREADS) k_leaves_colmajor(const u64 *__restrict__ data, size_t cols, size_t L, u64 *__restrict__ digests) {
    size_t leaf = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (leaf >= L) return;
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
    if (cols <= 4) {
#pragma unroll
        for (int c = 0; c < 4; c++) if ((size_t)c < cols) s[c] = data[c * L + leaf];
    } else {
        size_t c = 0;
        for (; c + 8 <= cols; c += 8) {
#pragma unroll
            for (int j = 0; j < 8; j++) s[j] = __ldg(data + (c + j) * L + leaf);
            poseidon_permute(s);
        }
        if (c < cols) {
#pragma unroll
            for (int j = 0; j < 8; j++) if (c + j < cols) s[j] = __ldg(data + (c + j) * L + leaf);
            poseidon_permute(s);
        }
    }
    ulonglong2 *o = reinterpret_cast<ulonglong2 *>(digests + 4 * leaf);
    o[0] = make_ulonglong2(s[0], s[1]); o[1] = make_ulonglong2(s[2], s[3]);
}

__global__ void __launch_bounds__(MK_THREADS) k_leaves_rowmajor(const u64 *__restrict__ data, size_t width, size_t L, u64 *__restrict__ digests) {
    size_t leaf = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (leaf >= L) return;
    const u64 *row = data + leaf * width;
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
    if (width <= 4) {
#pragma unroll
        for (int c = 0; c < 4; c++) if ((size_t)c < width) s[c] = row[c];
    } else {
        for (size_t c = 0; c < width; c += 8) {
#pragma unroll
            for (int j = 0; j < 8; j++) if (c + j < width) s[j] = row[c + j];
            poseidon_permute(s);
        }
    }
    ulonglong2 *o = reinterpret_cast<ulonglong2 *>(digests + 4 * leaf);
    o[0] = make_ulonglong2(s[0], s[1]); o[1] = make_ulonglong2(s[2], s[3]);
}

__global__ void __launch_bounds__(MK_THREADS) k_merkle_level(const u64 *__restrict__ in, u64 *__restrict__ out, size_t n_out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(in + 8 * i);
    ulonglong2 a = p[0], b = p[1], c = p[2], d = p[3];
    u64 s[12] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y, 0, 0, 0, 0};
    poseidon_permute(s);
    ulonglong2 *o = reinterpret_cast<ulonglong2 *>(out + 4 * i);
    o[0] = make_ulonglong2(s[0], s[1]); o[1] = make_ulonglong2(s[2], s[3]);
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
