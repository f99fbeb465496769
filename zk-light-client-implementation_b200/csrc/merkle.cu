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
#ifdef MK_MINB              /* experiments: force more resident blocks per SM (default: ptxas picks 80 registers = 6 blocks, no spills) */
#define MK_BOUNDS __launch_bounds__(MK_THREADS, MK_MINB)
#else
#define MK_BOUNDS __launch_bounds__(MK_THREADS)
#endif

// state of thread t lives in shared memory at sm[i * MK_THREADS + t], i < 12
#define ST(i) sm[(i) * MK_THREADS + threadIdx.x]
__global__ void MK_BOUNDS k_leaves_colmajor(const u64 *__restrict__ data, size_t cols, size_t L, u64 *__restrict__ digests) {
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

// ---------------------------------------------------------------------------------------------------------------
// Latency-oriented variant for the small launches (upper Merkle levels of every tree, FRI layer trees of small
// proofs): FOUR lanes cooperate on one permutation -- lane q holds state words 3q..3q+2, the S-box layer runs 3-wide per
// lane, the MDS layer gathers the 12 words with quad shuffles and each lane produces its 3 outputs.  A thread-per-
// permutation kernel needs ~40 us per tree level when the level cannot fill the machine (one warp walks 32.6 k
// dependent-ish instructions); the quad version cuts the per-level latency several times.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void quad_mds(u64 s[3], int q, int quad_base) {
    const u32 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    // gather the state already rotated by 3q words, so that every lane runs the same (divergence-free) dot products:
    // rot[j] = state[(j + 3q) % 12]
    u32 lo[12], hi[12];
#pragma unroll
    for (int d = 0; d < 4; d++) {
        const int src = quad_base + ((q + d) & 3);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lo[3 * d + k] = __shfl_sync(0xffffffffu, (u32)s[k], src);
            hi[3 * d + k] = __shfl_sync(0xffffffffu, (u32)(s[k] >> 32), src);
        }
    }
    const u32 diag = q == 0 ? 8u : 0u;                  // MDS diagonal: only out[0] += 8 * state[0]
#pragma unroll
    for (int k = 0; k < 3; k++) {
        u64 al = 0, ah = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) { al += (u64)lo[(i + k) % 12] * C[i]; ah += (u64)hi[(i + k) % 12] * C[i]; }
        if (k == 0) { al += (u64)lo[0] * diag; ah += (u64)hi[0] * diag; }
        s[k] = poseidon_reduce96(al, ah);
    }
}
// state words of this lane in s[3]; all 32 lanes of the warp must call (quads may hold dummy states)
__device__ __noinline__ void poseidon_permute_quad(u64 s[3], int q, int quad_base) {
    for (int r = 0; r < 30; r++) {
        const bool full = (r < 4) || (r >= 26);
        const u64 *rc = d_POSEIDON_RC + 12 * r + 3 * q;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            u64 c = rc[k];
            u64 x = s[k] + c;
            if (x < c) x += GL_EPS;
            if (full || (q == 0 && k == 0)) x = poseidon_sbox(x);
            s[k] = x;
        }
        quad_mds(s, q, quad_base);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) s[k] = gl_canon(s[k]);
}

__global__ void __launch_bounds__(128) k_merkle_level_quad(const u64 *__restrict__ in, u64 *__restrict__ out, size_t n_out) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t node = t >> 2; const int q = (int)(t & 3), quad_base = (int)(threadIdx.x & 31 & ~3);
    const bool active = node < n_out;
    if (!active) node = n_out - 1;
    const u64 *p = in + 8 * node;
    u64 s[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { int e = 3 * q + k; s[k] = e < 8 ? p[e] : 0; }
    poseidon_permute_quad(s, q, quad_base);
    if (active) {
        if (q == 0) { out[4 * node] = s[0]; out[4 * node + 1] = s[1]; out[4 * node + 2] = s[2]; }
        if (q == 1) out[4 * node + 3] = s[0];
    }
}
// row-major leaves (FRI layers): width > 4 words, sponge with rate 8
__global__ void __launch_bounds__(128) k_leaves_rowmajor_quad(const u64 *__restrict__ data, size_t width, size_t L, u64 *__restrict__ digests) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t leaf = t >> 2; const int q = (int)(t & 3), quad_base = (int)(threadIdx.x & 31 & ~3);
    const bool active = leaf < L;
    if (!active) leaf = L - 1;
    const u64 *row = data + leaf * width;
    u64 s[3] = {0, 0, 0};
    for (size_t c = 0; c < width; c += 8) {
#pragma unroll
        for (int k = 0; k < 3; k++) { int e = 3 * q + k; if (e < 8 && c + e < width) s[k] = row[c + e]; }
        poseidon_permute_quad(s, q, quad_base);
    }
    if (active) {
        if (q == 0) { digests[4 * leaf] = s[0]; digests[4 * leaf + 1] = s[1]; digests[4 * leaf + 2] = s[2]; }
        if (q == 1) digests[4 * leaf + 3] = s[0];
    }
}
// ---------------------------------------------------------------------------------------------------------------
// TWELVE lanes per permutation (one state word per lane), for launches so small that only the latency of ONE permutation
// matters (ncu on the quad form: ~28 us per tree level, ~470 dependent instructions per round per lane; here ~185: the lane
// adds its round constant, raises its own word to the 7th power, and builds its MDS row from the other eleven words with
// 22 shuffles and 24 wide multiply-accumulates).  A warp carries two permutations (lanes 0-11, 12-23); lanes 24-31 run the
// same instructions on dummy data and never store.  Used for every launch of <= X12_MAX_NODES permutations (a
// recursion-sized proof spends 0.26 ms less in its Merkle launches: profiles/phase_rec13_r01_v9_x12_optin.json).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 poseidon_permute_x12(u64 x, const int w, const int base) {
    const u32 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    for (int r = 0; r < 30; r++) {
        const bool full = (r < 4) || (r >= 26);
        const u64 c = d_POSEIDON_RC[12 * r + w];
        u64 v = x + c;
        if (v < c) v += GL_EPS;
        if (full || w == 0) v = poseidon_sbox(v);
        const u32 lo = (u32)v, hi = (u32)(v >> 32);
        u64 al = 0, ah = 0;
        int src = w;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            const u32 nl = __shfl_sync(0xffffffffu, lo, base + src), nh = __shfl_sync(0xffffffffu, hi, base + src);
            al += (u64)nl * C[i]; ah += (u64)nh * C[i];
            if (++src == 12) src = 0;
        }
        if (w == 0) { al += (u64)lo * 8u; ah += (u64)hi * 8u; }      // MDS diagonal
        x = poseidon_reduce96(al, ah);
    }
    return gl_canon(x);
}
#define X12_PER_BLOCK 8            /* 4 warps x 2 permutations */
__global__ void __launch_bounds__(128) k_merkle_level_x12(const u64 *__restrict__ in, u64 *__restrict__ out, size_t n_out) {
    const int lane = threadIdx.x & 31, g = lane / 12, w = lane - 12 * g, base = 12 * g;
    size_t node = (size_t)blockIdx.x * X12_PER_BLOCK + (threadIdx.x >> 5) * 2 + g;
    const bool active = g < 2 && node < n_out;
    if (node >= n_out) node = n_out - 1;
    u64 x = (w < 8) ? in[8 * node + w] : 0;
    x = poseidon_permute_x12(x, w, base);
    if (active && w < 4) out[4 * node + w] = x;
}
__global__ void __launch_bounds__(128) k_leaves_rowmajor_x12(const u64 *__restrict__ data, size_t width, size_t L, u64 *__restrict__ digests) {
    const int lane = threadIdx.x & 31, g = lane / 12, w = lane - 12 * g, base = 12 * g;
    size_t leaf = (size_t)blockIdx.x * X12_PER_BLOCK + (threadIdx.x >> 5) * 2 + g;
    const bool active = g < 2 && leaf < L;
    if (leaf >= L) leaf = L - 1;
    const u64 *row = data + leaf * width;
    u64 x = 0;
    for (size_t c = 0; c < width; c += 8) {
        if (w < 8 && c + w < width) x = row[c + w];
        x = poseidon_permute_x12(x, w, base);
    }
    if (active && w < 4) digests[4 * leaf + w] = x;
}
#define X12_MAX_NODES 4096          /* above this the quad form's throughput wins */

#define QUAD_MAX_NODES (size_t(148) * 16 * 32 / 4)      /* below ~19 k permutations a launch cannot fill the machine thread-per-permutation (2^15 on the quad form measured slower: profiles/phase_rec13_r02_v8.json) */

void merkle_leaves_colmajor(const u64 *data, size_t cols, int log_L, u64 *digests, cudaStream_t st) {
    size_t L = size_t(1) << log_L;
    k_leaves_colmajor<<<(unsigned)((L + MK_THREADS - 1) / MK_THREADS), MK_THREADS, 0, st>>>(data, cols, L, digests);
    LAUNCHED(1);
}
void merkle_leaves_rowmajor(const u64 *data, size_t width, int log_L, u64 *digests, cudaStream_t st) {
    size_t L = size_t(1) << log_L;
    if (width > 4 && L <= X12_MAX_NODES) k_leaves_rowmajor_x12<<<(unsigned)((L + X12_PER_BLOCK - 1) / X12_PER_BLOCK), 128, 0, st>>>(data, width, L, digests);
    else if (width > 4 && L <= QUAD_MAX_NODES) k_leaves_rowmajor_quad<<<(unsigned)((4 * L + 127) / 128), 128, 0, st>>>(data, width, L, digests);
    else k_leaves_rowmajor<<<(unsigned)((L + MK_THREADS - 1) / MK_THREADS), MK_THREADS, 0, st>>>(data, width, L, digests);
    LAUNCHED(1);
}
void merkle_build(u64 *digests, int log_L, int cap_height, cudaStream_t st) {
    u64 *cur = digests;
    for (int lv = log_L; lv > cap_height; lv--) {
        size_t n = size_t(1) << lv;
        u64 *nxt = cur + 4 * n;
        if (n / 2 <= X12_MAX_NODES) k_merkle_level_x12<<<(unsigned)((n / 2 + X12_PER_BLOCK - 1) / X12_PER_BLOCK), 128, 0, st>>>(cur, nxt, n / 2);
        else if (n / 2 <= QUAD_MAX_NODES) k_merkle_level_quad<<<(unsigned)((4 * (n / 2) + 127) / 128), 128, 0, st>>>(cur, nxt, n / 2);
        else k_merkle_level<<<(unsigned)((n / 2 + MK_THREADS - 1) / MK_THREADS), MK_THREADS, 0, st>>>(cur, nxt, n / 2);
        LAUNCHED(1);
        cur = nxt;
    }
}
