// Batched Goldilocks DIF NTT / coset LDE passes staged through shared memory with the bulk asynchronous copy engine (TMA:
// cp.async.bulk global->shared completing on an mbarrier, shared->global in bulk groups; SASS UBLKCP), for the transform
// sizes the prover spends its NTT time on (2^16 .. 2^20 points: witness / quotient-chunk iNTT + x8 coset LDE at
// degree_bits 13..17, the 2^20 quotient iNTT).  Replaces, for those sizes, the load / store phases and the per-butterfly
// twiddle gathers of ntt.cu's k_ntt_dif_pass, which the round-1 profile showed to be half of its instruction stream
// (profiles/ncu_k_ntt_dif_pass_r01_v3_pass0.json: 76 instructions per butterfly, 40 of them arithmetic).
// Upstream routine replaced: plonky2_field::fft::{fft_dispatch, ifft}, PolynomialValues::lde / coset_fft as used by
// PolynomialBatch::from_values / from_coeffs (loop A of section 8).
//
// Structure (index math in ntt2_core.cuh, checked on the host by tests/ntt2_emulation.cu):
//   * a transform of 2^L points, 16 <= L <= 20, is two passes over 4096-element tiles: 8 layers on strided tiles
//     (256 x 16-element rows, every row one 128-byte bulk copy), then L - 8 layers on contiguous tiles;
//   * blocks are persistent over (tile position, item) work units in position-major order, so the twiddles of a round
//     live in registers / a 2 KB shared table for hundreds of consecutive units instead of being gathered per butterfly;
//   * a round = every thread pulls one radix-16 group (16 u64) into registers, runs 4 butterfly layers, writes back;
//   * the unit's input rows are fetched by 256 bulk copies (one per thread) that complete on an mbarrier while the
//     previous unit is still in its later rounds; results leave by bulk stores from one of two work buffers, so a store
//     drains while the next unit computes in the other buffer; two blocks per SM overlap the rest;
//   * coset LDE (first pass): the coefficient tile arrives in bit-reversed storage order as 16-element runs that are
//     exactly the radix-16 groups of the first round; the coset scale factors of a run are 128 contiguous bytes of the
//     table (same storage order) and are read with 128-bit loads by the thread that owns the run.
#include "b200_internal.h"
#include "ntt2_core.cuh"
#include <algorithm>
#include <atomic>
#include <string.h>

struct N2Args {
    const u64 *in; u64 *out; const u64 *tw; const u64 *scale;
    u64 in_stride, out_stride;      // elements between consecutive columns (first pass) / items
    N2Pass p;
    int rate_bits;                  // first pass of an LDE: item = (column << rate_bits) | coset
    u64 out_mul;                    // applied in the last round of the last pass (1 = none)
    u32 items, tiles_per_item;
};

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src_smem, u32 bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// MODE 0: strided pass, natural input; 1: strided first pass of an LDE (bit-reversed input + coset scale); 2: contiguous last pass
template <int MODE>
__global__ void __launch_bounds__(N2_THREADS, 2) k_ntt2_pass(const N2Args a) {
    extern __shared__ __align__(128) u64 sm[];
    u64 *IN = sm, *W0 = sm + N2_BUF_WORDS, *W1 = sm + 2 * N2_BUF_WORDS;        // input rows; two work buffers (the bulk store of unit w drains while unit w + 1 computes in the other)
    u64 *twB = sm + 3 * N2_BUF_WORDS;                                          // [16][16] round-table (15 used per row)
    u64 *bar = twB + 256 + 16;
    const u32 tid = threadIdx.x;
    const N2Pass p = a.p;
    const u64 W = (u64)a.items * a.tiles_per_item;
    const u64 w0 = W * blockIdx.x / gridDim.x, w1 = W * (blockIdx.x + 1) / gridDim.x;
    if (w0 >= w1) return;
    if (tid == 0) mbar_init(bar, N2_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const u64 N = u64(1) << p.L;
    // item -> base pointers
    auto in_base = [&](u32 item) -> const u64 * {
        if (MODE == 1) return a.in + (u64)(item >> a.rate_bits) * a.in_stride;
        return a.in + (u64)item * a.in_stride;
    };
    auto out_base = [&](u32 item) -> u64 * {
        if (MODE == 1) return a.out + (u64)(item >> a.rate_bits) * a.out_stride + (u64)bitrev32(item & ((1u << a.rate_bits) - 1), a.rate_bits) * N;
        return a.out + (u64)item * a.out_stride;
    };
    auto issue_load = [&](u64 w) {
        const u32 pos = (u32)(w / a.items), item = (u32)(w % a.items);
        const u64 off = MODE == 1 ? n2_row_global_bitrev(p, pos, tid) : n2_row_global(p, pos, tid);
        mbar_arrive_expect_tx(bar, 128u);
        bulk_g2s(IN + tid * N2_ROW_WORDS, in_base(item) + off, 128, bar);
    };

    // twiddles of the first round of a unit live in registers; they depend on the tile position only for the strided passes
    u64 twsA[15];
#pragma unroll
    for (int i = 0; i < 15; i++) twsA[i] = 0;
    u32 cur_pos = 0xFFFFFFFFu;
    if (MODE == 2) {                                             // contiguous pass: nothing depends on the tile position
        n2_build_twsA<2>(p, a.tw, 0, tid, twsA);
        if (tid < 240) twB[(tid / 15) * 16 + tid % 15] = n2_twB_entry<2>(p, a.tw, 0, tid / 15, tid % 15);
        __syncthreads();
    }

    issue_load(w0);
    u32 parity = 0;
    for (u64 w = w0; w < w1; w++) {
        const u32 pos = (u32)(w / a.items), item = (u32)(w % a.items);
        u64 *WORK = ((w - w0) & 1) ? W1 : W0;
        bulk_wait_read1();                                       // my row of the unit before the previous one has left this work buffer
        if (MODE != 2 && pos != cur_pos) {                       // block-uniform
            cur_pos = pos;
            n2_build_twsA<MODE>(p, a.tw, pos, tid, twsA);
            __syncthreads();                                     // the previous unit's rounds have read twB
            if (tid < 240) twB[(tid / 15) * 16 + tid % 15] = n2_twB_entry<MODE>(p, a.tw, pos, tid / 15, tid % 15);
        }
        mbar_wait(bar, parity); parity ^= 1;
        __syncthreads();                                         // everyone's older store has been read; twB visible
        const u64 *sc_row = (MODE == 1 && a.scale) ? a.scale + (u64)(item & ((1u << a.rate_bits) - 1)) * N + n2_row_global_bitrev(p, pos, tid) : nullptr;
        n2_first_round<MODE>(p, IN, sc_row, WORK, tid, twsA, twB);     // IN (x coset scale) -> WORK
        __syncthreads();
        if (w + 1 < w1) issue_load(w + 1);                       // IN / SC are free: fetch the next unit behind the remaining rounds
        if (n2_second_round<MODE>(p, WORK, tid, twB, a.out_mul)) { __syncthreads(); n2_third_round(p, WORK, tid, a.out_mul); }
        fence_async_smem();
        __syncthreads();
        bulk_s2g(out_base(item) + n2_row_global(p, pos, tid), WORK + tid * N2_ROW_WORDS, 128);
        bulk_commit();
    }
    bulk_wait0();
}

static size_t n2_smem_bytes(int) { return (3 * N2_BUF_WORDS + 256 + 16 + 2) * sizeof(u64); }

bool ntt2_supported(int L, size_t in_stride, size_t out_stride, int rate_bits) {
    if (L < 16 || L > 20) return false;
    if ((in_stride & 1) || (out_stride & 1)) return false;      // 16-byte aligned rows
    (void)rate_bits;
    return true;
}

// Same contract as ntt_dif (b200_internal.h) for 16 <= L <= 20.
void ntt2_dif(const NttTables &T, const u64 *in, size_t in_stride, bool in_bitrev, u64 *out, size_t out_stride,
              int L, size_t items, int rate_bits, const u64 *scale, bool inverse, u64 out_mul, cudaStream_t st) {
    // per-device one-time set-up (several contexts / host threads may get here concurrently: the calls are idempotent)
    static std::atomic<unsigned long long> dev_ready{0};
    static int sms_of[64];
    int dev = 0; CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) throw CudaErr{cudaErrorInvalidDevice, __FILE__, __LINE__};
    if (!((dev_ready.load() >> dev) & 1)) {
        int n_sm = 0; CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        sms_of[dev] = n_sm > 0 ? n_sm : 148;
        CUDA_TRY(cudaFuncSetAttribute(k_ntt2_pass<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)n2_smem_bytes(0)));
        CUDA_TRY(cudaFuncSetAttribute(k_ntt2_pass<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)n2_smem_bytes(1)));
        CUDA_TRY(cudaFuncSetAttribute(k_ntt2_pass<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)n2_smem_bytes(2)));
        dev_ready.fetch_or(1ULL << dev);
    }
    const int sms = sms_of[dev];
    const u64 *tw = inverse ? T.inv : T.fwd;
    // pass 1: layers [0, 8) on strided tiles
    N2Args a; memset(&a, 0, sizeof a);
    a.in = in; a.out = out; a.tw = tw; a.scale = scale; a.in_stride = in_stride; a.out_stride = out_stride;
    a.p.L = L; a.p.l_begin = 0; a.p.B = 8; a.p.strided = 1; a.p.in_bitrev = in_bitrev ? 1 : 0; a.p.logS = L - 8; a.p.inverse = inverse ? 1 : 0;
    a.rate_bits = rate_bits; a.out_mul = 1; a.items = (u32)items; a.tiles_per_item = 1u << (L - 12);
    const bool lde_first = in_bitrev;                            // bit-reversed storage in (with or without a coset scale)
    if (!lde_first && rate_bits) throw CudaErr{cudaErrorInvalidValue, __FILE__, __LINE__};
    if (!lde_first && scale) throw CudaErr{cudaErrorInvalidValue, __FILE__, __LINE__};
    unsigned grid = (unsigned)std::min<u64>((u64)a.items * a.tiles_per_item, (u64)2 * sms);
    if (lde_first) k_ntt2_pass<1><<<grid, N2_THREADS, n2_smem_bytes(1), st>>>(a);
    else k_ntt2_pass<0><<<grid, N2_THREADS, n2_smem_bytes(0), st>>>(a);
    LAUNCHED(1);
    // pass 2: layers [8, L) on contiguous tiles, in place in `out`; after an LDE first pass the cosets of a column are
    // consecutive 2^L blocks of `out`
    N2Args b = a;
    b.in = out; b.scale = nullptr; b.rate_bits = 0; b.out_mul = out_mul;
    if (rate_bits) {
        if (out_stride != (size_t(1) << (L + rate_bits))) throw CudaErr{cudaErrorInvalidValue, __FILE__, __LINE__};
        b.in_stride = b.out_stride = u64(1) << L;
    } else b.in_stride = b.out_stride = out_stride;
    b.p.l_begin = 8; b.p.B = L - 8; b.p.strided = 0; b.p.in_bitrev = 0; b.p.logS = 0;
    k_ntt2_pass<2><<<grid, N2_THREADS, n2_smem_bytes(2), st>>>(b);
    LAUNCHED(1);
}
