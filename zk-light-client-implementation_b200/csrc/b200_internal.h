// Internal C++ declarations shared by the .cu translation units of libb200plonky2.so.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "gl.cuh"

#define B200_OK 0
#define B200_ERR_BAD_ARG (-1)
#define B200_ERR_CUDA (-2)
#define B200_ERR_UNSATISFIED (-3)
#define B200_ERR_OOM (-4)
#define B200_ERR_UNSUPPORTED (-5)
#define B200_ERR_INVALID_PROOF (-6)

struct CudaErr { cudaError_t e; const char *file; int line; };
#define CUDA_TRY(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) throw CudaErr{_e, __FILE__, __LINE__}; } while (0)

#include <atomic>
extern std::atomic<unsigned long long> g_b200_launches;     // kernels launched by this library (process-wide)
#define LAUNCHED(n) do { g_b200_launches += (n); CUDA_TRY(cudaGetLastError()); } while (0)

struct DevBuf {   // RAII device buffer
    void *p = nullptr; size_t bytes = 0;
    DevBuf() {}
    explicit DevBuf(size_t b) { alloc(b); }
    void alloc(size_t b) { release(); bytes = b; if (b) { cudaError_t e = cudaMalloc(&p, b); if (e != cudaSuccess) throw CudaErr{e, __FILE__, __LINE__}; } }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    ~DevBuf() { release(); }
    DevBuf(const DevBuf &) = delete; DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept { p = o.p; bytes = o.bytes; o.p = nullptr; o.bytes = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept { release(); p = o.p; bytes = o.bytes; o.p = nullptr; o.bytes = 0; return *this; }
    u64 *u() const { return (u64 *)p; }
};

// ---- per-device tables (twiddles etc.), owned by the context ----
struct NttTables {
    int max_log = 0;
    u64 *fwd = nullptr;   // fwd[h + j] = w_{2h}^j, h = 2^s, j < h            (size 2^max_log)
    u64 *inv = nullptr;   // inv[h + j] = w_{2h}^{-j}
};
void ntt_tables_ensure(NttTables &T, int log_n, cudaStream_t st);
void ntt_tables_free(NttTables &T);

// DIF transform of `items` polynomials of size 2^L.
//   in  : item it -> in  + (it >> rate_bits) * in_stride      (natural order, or bit-reversed storage if in_bitrev)
//   out : item it -> out + (it >> rate_bits) * out_stride + bitrev(it & (2^rate_bits-1)) * 2^L
//   scale (optional): element at storage position p of item it is multiplied by scale[(it & mask) * 2^L + p] at load
//   result stored in bit-reversed order (standard in-place DIF), multiplied by out_mul (1 = none)
void ntt_dif(const NttTables &T, const u64 *in, size_t in_stride, bool in_bitrev, u64 *out, size_t out_stride,
             int L, size_t items, int rate_bits, const u64 *scale, bool inverse, u64 out_mul, cudaStream_t st);
// bulk-copy (TMA) staged passes for 16 <= L <= 20 (ntt2.cu); ntt_dif dispatches to them
bool ntt2_supported(int L, size_t in_stride, size_t out_stride, int rate_bits);
void ntt2_dif(const NttTables &T, const u64 *in, size_t in_stride, bool in_bitrev, u64 *out, size_t out_stride,
              int L, size_t items, int rate_bits, const u64 *scale, bool inverse, u64 out_mul, cudaStream_t st);
void bitrev_permute(const u64 *in, u64 *out, int L, size_t items, cudaStream_t st);   // out[it][bitrev(i)] = in[it][i]
// powers table: out[r * n + p] = (shift * w_{n<<rate_bits}^r)^{bitrev_L(p)} * mul   (coset scaling for the LDE)
void lde_scale_table(u64 *out, int L, int rate_bits, u64 shift, cudaStream_t st);

// ---- Poseidon / Merkle ----
// digests layout: level 0 (L x 4 u64), level 1 (L/2 x 4), ..., cap level (2^cap_height x 4)
static inline size_t merkle_num_u64(int log_L, int cap_height) { return 4 * ((size_t(2) << log_L) - (size_t(1) << cap_height)); }
void merkle_leaves_colmajor(const u64 *data, size_t cols, int log_L, u64 *digests, cudaStream_t st);
void merkle_leaves_rowmajor(const u64 *data, size_t width, int log_L, u64 *digests, cudaStream_t st);
void merkle_build(u64 *digests, int log_L, int cap_height, cudaStream_t st);
// PoseidonBN128Hash variants (merkle_bn254.cu)
void merkle_leaves_colmajor_bn254(const u64 *data, size_t cols, int log_L, u64 *digests, cudaStream_t st);
void merkle_leaves_rowmajor_bn254(const u64 *data, size_t width, int log_L, u64 *digests, cudaStream_t st);
void merkle_build_bn254(u64 *digests, int log_L, int cap_height, cudaStream_t st);
enum { HASHER_GL = 0, HASHER_BN254 = 1 };
