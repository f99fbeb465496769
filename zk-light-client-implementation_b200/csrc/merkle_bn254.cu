// Merkle-cap commitment kernels for PoseidonBN128Hash (the final BN128 wrap proof, SURVEY.md section 8 row a5:
// near_bft_finality/src/bin/prove_block.rs:279-300 with crypto/plonky2_bn128/src/config.rs:132-200).
// Same data layout and digest storage as merkle.cu (4 u64 per digest = 32 LE bytes of the canonical Fr value).
#include "b200_internal.h"
#include "bn254.cuh"

#define BK_THREADS 64

__device__ __forceinline__ void absorb_and_permute(fr_t s[4], const u64 v[9], int m) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (3 * j < m) {
            u64 l[4] = {v[3 * j], (3 * j + 1 < m) ? v[3 * j + 1] : 0, (3 * j + 2 < m) ? v[3 * j + 2] : 0, 0};
            s[j + 1] = fr_from_canonical(l);
        }
    }
    bn254_permute(s);
}

__global__ void __launch_bounds__(BK_THREADS) k_leaves_colmajor_bn254(const u64 *__restrict__ data, size_t cols, size_t L, u64 *__restrict__ digests) {
    size_t leaf = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (leaf >= L) return;
    u64 out[4] = {0, 0, 0, 0};
    if (cols <= 3) {
        for (size_t c = 0; c < cols; c++) out[c] = data[c * L + leaf];
    } else {
        fr_t s[4];
        for (int i = 0; i < 4; i++) for (int k = 0; k < 4; k++) s[i].v[k] = 0;
        for (size_t c = 0; c < cols; c += 9) {
            u64 v[9]; int m = (int)((cols - c < 9) ? cols - c : 9);
#pragma unroll
            for (int j = 0; j < 9; j++) v[j] = (j < m) ? __ldg(data + (c + j) * L + leaf) : 0;
            absorb_and_permute(s, v, m);
        }
        fr_to_canonical(s[0], out);
    }
    for (int k = 0; k < 4; k++) digests[4 * leaf + k] = out[k];
}

__global__ void __launch_bounds__(BK_THREADS) k_leaves_rowmajor_bn254(const u64 *__restrict__ data, size_t width, size_t L, u64 *__restrict__ digests) {
    size_t leaf = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (leaf >= L) return;
    const u64 *row = data + leaf * width;
    u64 out[4] = {0, 0, 0, 0};
    if (width <= 3) {
        for (size_t c = 0; c < width; c++) out[c] = row[c];
    } else {
        fr_t s[4];
        for (int i = 0; i < 4; i++) for (int k = 0; k < 4; k++) s[i].v[k] = 0;
        for (size_t c = 0; c < width; c += 9) {
            u64 v[9]; int m = (int)((width - c < 9) ? width - c : 9);
#pragma unroll
            for (int j = 0; j < 9; j++) v[j] = (j < m) ? row[c + j] : 0;
            absorb_and_permute(s, v, m);
        }
        fr_to_canonical(s[0], out);
    }
    for (int k = 0; k < 4; k++) digests[4 * leaf + k] = out[k];
}

__global__ void __launch_bounds__(BK_THREADS) k_merkle_level_bn254(const u64 *__restrict__ in, u64 *__restrict__ out, size_t n_out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    u64 l[4], r[4], o[4];
    for (int k = 0; k < 4; k++) { l[k] = in[8 * i + k]; r[k] = in[8 * i + 4 + k]; }
    bn254_two_to_one(l, r, o);
    for (int k = 0; k < 4; k++) out[4 * i + k] = o[k];
}

void merkle_leaves_colmajor_bn254(const u64 *data, size_t cols, int log_L, u64 *digests, cudaStream_t st) {
    size_t L = size_t(1) << log_L;
    k_leaves_colmajor_bn254<<<(unsigned)((L + BK_THREADS - 1) / BK_THREADS), BK_THREADS, 0, st>>>(data, cols, L, digests);
    LAUNCHED(1);
}
void merkle_leaves_rowmajor_bn254(const u64 *data, size_t width, int log_L, u64 *digests, cudaStream_t st) {
    size_t L = size_t(1) << log_L;
    k_leaves_rowmajor_bn254<<<(unsigned)((L + BK_THREADS - 1) / BK_THREADS), BK_THREADS, 0, st>>>(data, width, L, digests);
    LAUNCHED(1);
}
void merkle_build_bn254(u64 *digests, int log_L, int cap_height, cudaStream_t st) {
    u64 *cur = digests;
    for (int lv = log_L; lv > cap_height; lv--) {
        size_t n = size_t(1) << lv;
        u64 *nxt = cur + 4 * n;
        k_merkle_level_bn254<<<(unsigned)((n / 2 + BK_THREADS - 1) / BK_THREADS), BK_THREADS, 0, st>>>(cur, nxt, n / 2);
        LAUNCHED(1);
        cur = nxt;
    }
}
