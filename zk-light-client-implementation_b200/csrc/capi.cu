// extern "C" surface of libb200plonky2.so (see include/b200_plonky2.h): context management and the
// inner-loop entry points.  b200_circuit_upload / b200_prove live in prover.cu.
#include "../../include/b200_plonky2.h"
#include "b200_internal.h"
#include "ctx.h"
#include <stdlib.h>
#include <string.h>

std::atomic<unsigned long long> g_b200_launches{0};

int ctx_fail(b200_ctx *ctx, int code, const std::string &msg) { if (ctx) ctx->err = msg; return code; }

#define API_BEGIN(ctx) if (!(ctx)) return B200_ERR_BAD_ARG; try { cudaSetDevice((ctx)->device);
#define API_END(ctx) } catch (const CudaErr &e) { \
        char buf[256]; snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d", (int)e.e, cudaGetErrorString(e.e), e.file, e.line); \
        cudaGetLastError(); \
        return ctx_fail(ctx, e.e == cudaErrorMemoryAllocation ? B200_ERR_OOM : B200_ERR_CUDA, buf); \
    } catch (const ApiErr &e) { return ctx_fail(ctx, e.code, e.msg); \
    } catch (const std::bad_alloc &) { return ctx_fail(ctx, B200_ERR_OOM, "host allocation failed"); \
    } catch (const std::exception &e) { return ctx_fail(ctx, B200_ERR_BAD_ARG, e.what()); }

extern "C" {

int b200_ctx_create(int device, b200_ctx **out) {
    if (!out) return B200_ERR_BAD_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) { cudaGetLastError(); return B200_ERR_CUDA; }
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return B200_ERR_CUDA; }
    b200_ctx *c = new (std::nothrow) b200_ctx();
    if (!c) return B200_ERR_OOM;
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; cudaGetLastError(); return B200_ERR_CUDA; }
    c->own_stream = true;
    *out = c;
    return B200_OK;
}

void b200_ctx_destroy(b200_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx_drop_circuit_cache(ctx);
    ntt_tables_free(ctx->tw);
    for (auto &kv : ctx->lde_scale) cudaFree(kv.second);
    if (ctx->arena) cudaFree(ctx->arena);
    for (int i = 0; i < 2; i++) if (ctx->fr_tw[i]) cudaFree(ctx->fr_tw[i]);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int b200_ctx_set_stream(b200_ctx *ctx, void *cuda_stream) {
    API_BEGIN(ctx)
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) { cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
    if (cuda_stream) ctx->stream = (cudaStream_t)cuda_stream;
    else { CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
    return B200_OK;
    API_END(ctx)
}

int b200_ctx_synchronize(b200_ctx *ctx) {
    API_BEGIN(ctx)
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
    API_END(ctx)
}

const char *b200_last_error(const b200_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void b200_free(void *p) { free(p); }
int b200_host_alloc(b200_ctx *ctx, size_t bytes, void **out) {
    API_BEGIN(ctx)
    if (!out) throw ApiErr{B200_ERR_BAD_ARG, "b200_host_alloc: null out"};
    *out = nullptr;
    if (cudaMallocHost(out, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); throw ApiErr{B200_ERR_OOM, "b200_host_alloc: cudaMallocHost failed"}; }
    return B200_OK;
    API_END(ctx)
}
void b200_host_free(void *p) { if (p) cudaFreeHost(p); }
uint64_t b200_kernel_launches(const b200_ctx *) { return g_b200_launches.load(); }

int b200_ctx_enable_timing(b200_ctx *ctx, int enable) {
    API_BEGIN(ctx)
    ctx->timing = enable != 0;
    return B200_OK;
    API_END(ctx)
}

// JSON object {"phase": [total_ms, launches], ...} accumulated since the last call; clears the record.
const char *b200_ctx_timings(b200_ctx *ctx) {
    if (!ctx) return "{}";
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    std::map<std::string, std::pair<double, int>> acc;
    std::vector<std::string> order;
    for (auto &e : ctx->timing_events) {
        float ms = 0; cudaEventElapsedTime(&ms, e.second.first, e.second.second);
        if (!acc.count(e.first)) order.push_back(e.first);
        acc[e.first].first += ms; acc[e.first].second += 1;
        cudaEventDestroy(e.second.first); cudaEventDestroy(e.second.second);
    }
    ctx->timing_events.clear();
    std::string j = "{";
    for (size_t i = 0; i < order.size(); i++) {
        char buf[160]; snprintf(buf, sizeof buf, "%s\"%s\": [%.4f, %d]", i ? ", " : "", order[i].c_str(), acc[order[i]].first, acc[order[i]].second);
        j += buf;
    }
    if (ctx->host_proofs) {
        char buf[200]; snprintf(buf, sizeof buf, "%s\"_host_total\": [%.4f, %d], \"_host_sync_wait\": [%.4f, %d]", order.empty() ? "" : ", ",
                                ctx->host_total_ms, ctx->host_proofs, ctx->host_wait_ms, ctx->host_syncs);
        j += buf;
        for (size_t i = 0; i < ctx->host_seg_ms.size(); i++) { snprintf(buf, sizeof buf, ", \"_host_before_sync_%02zu\": [%.4f, %d]", i, ctx->host_seg_ms[i], ctx->host_proofs); j += buf; }
        ctx->host_seg_ms.clear();
        ctx->host_total_ms = ctx->host_wait_ms = 0; ctx->host_proofs = ctx->host_syncs = 0;
    }
    j += "}";
    ctx->timing_json = j;
    return ctx->timing_json.c_str();
}
size_t b200_merkle_num_u64(uint32_t log_L, uint32_t cap_height) { return merkle_num_u64((int)log_L, (int)cap_height); }

// ---------------- NTT ----------------
int b200_ntt_batch_dev(b200_ctx *ctx, uint64_t *d_values, size_t cols, uint32_t log_n, int inverse, uint64_t *d_scratch) {
    API_BEGIN(ctx)
    if (!d_values || !d_scratch || cols == 0 || log_n == 0 || log_n > 30) throw ApiErr{B200_ERR_BAD_ARG, "b200_ntt_batch: bad arguments"};
    ntt_tables_ensure(ctx->tw, (int)log_n, ctx->stream);
    size_t n = size_t(1) << log_n;
    u64 mul = inverse ? gl_inv((u64)n % GL_P) : 1;
    ntt_dif(ctx->tw, d_values, n, false, d_scratch, n, (int)log_n, cols, 0, nullptr, inverse != 0, mul, ctx->stream);
    bitrev_permute(d_scratch, d_values, (int)log_n, cols, ctx->stream);
    return B200_OK;
    API_END(ctx)
}

int b200_ntt_batch(b200_ctx *ctx, uint64_t *values, size_t cols, uint32_t log_n, int inverse) {
    API_BEGIN(ctx)
    if (!values || cols == 0 || log_n == 0 || log_n > 30) throw ApiErr{B200_ERR_BAD_ARG, "b200_ntt_batch: bad arguments"};
    size_t bytes = (cols << log_n) * sizeof(u64);
    DevBuf a(bytes), b(bytes);
    CUDA_TRY(cudaMemcpyAsync(a.p, values, bytes, cudaMemcpyHostToDevice, ctx->stream));
    int rc = b200_ntt_batch_dev(ctx, a.u(), cols, log_n, inverse, b.u());
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(values, a.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
    API_END(ctx)
}

// ---------------- LDE + commit ----------------
int b200_lde_commit_dev(b200_ctx *ctx, const uint64_t *d_values, size_t cols, uint32_t log_n, uint32_t rate_bits,
                        uint32_t cap_height, uint64_t *d_coeffs_br, uint64_t *d_lde, uint64_t *d_digests) {
    API_BEGIN(ctx)
    if (!d_values || !d_coeffs_br || !d_lde || !d_digests || cols == 0 || log_n == 0 || log_n + rate_bits > 30 || rate_bits > 4 ||
        cap_height > log_n + rate_bits)
        throw ApiErr{B200_ERR_BAD_ARG, "b200_lde_commit: bad arguments"};
    commit_values(ctx, d_values, cols, (int)log_n, (int)rate_bits, (int)cap_height, d_coeffs_br, d_lde, d_digests);
    return B200_OK;
    API_END(ctx)
}

int b200_lde_commit(b200_ctx *ctx, const uint64_t *values, size_t cols, uint32_t log_n, uint32_t rate_bits,
                    uint32_t cap_height, uint64_t *cap_out, uint64_t *lde_out, uint64_t *coeffs_out) {
    API_BEGIN(ctx)
    if (!values || !cap_out || cols == 0) throw ApiErr{B200_ERR_BAD_ARG, "b200_lde_commit: bad arguments"};
    size_t n = size_t(1) << log_n, N = n << rate_bits;
    DevBuf v(cols * n * 8), c(cols * n * 8), lde(cols * N * 8), dg(merkle_num_u64((int)(log_n + rate_bits), (int)cap_height) * 8);
    CUDA_TRY(cudaMemcpyAsync(v.p, values, cols * n * 8, cudaMemcpyHostToDevice, ctx->stream));
    int rc = b200_lde_commit_dev(ctx, v.u(), cols, log_n, rate_bits, cap_height, c.u(), lde.u(), dg.u());
    if (rc) return rc;
    size_t cap_off = merkle_num_u64((int)(log_n + rate_bits), (int)cap_height) - (size_t(4) << cap_height);
    CUDA_TRY(cudaMemcpyAsync(cap_out, dg.u() + cap_off, (size_t(32) << cap_height), cudaMemcpyDeviceToHost, ctx->stream));
    if (lde_out) CUDA_TRY(cudaMemcpyAsync(lde_out, lde.p, cols * N * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (coeffs_out) {
        bitrev_permute(c.u(), v.u(), (int)log_n, cols, ctx->stream);      // br storage -> natural order
        CUDA_TRY(cudaMemcpyAsync(coeffs_out, v.p, cols * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
    API_END(ctx)
}

// ---------------- Merkle ----------------
int b200_poseidon_merkle_dev(b200_ctx *ctx, const uint64_t *d_data, size_t cols, uint32_t log_L, uint32_t cap_height, uint64_t *d_digests) {
    API_BEGIN(ctx)
    if (!d_data || !d_digests || cols == 0 || log_L > 30 || cap_height > log_L) throw ApiErr{B200_ERR_BAD_ARG, "b200_poseidon_merkle: bad arguments"};
    merkle_leaves_colmajor(d_data, cols, (int)log_L, d_digests, ctx->stream);
    merkle_build(d_digests, (int)log_L, (int)cap_height, ctx->stream);
    return B200_OK;
    API_END(ctx)
}

int b200_poseidon_merkle(b200_ctx *ctx, const uint64_t *data, size_t cols, uint32_t log_L, uint32_t cap_height,
                         uint64_t *cap_out, uint64_t *digests_out) {
    API_BEGIN(ctx)
    if (!data || !cap_out) throw ApiErr{B200_ERR_BAD_ARG, "b200_poseidon_merkle: bad arguments"};
    size_t L = size_t(1) << log_L, nd = merkle_num_u64((int)log_L, (int)cap_height);
    DevBuf d(cols * L * 8), dg(nd * 8);
    CUDA_TRY(cudaMemcpyAsync(d.p, data, cols * L * 8, cudaMemcpyHostToDevice, ctx->stream));
    int rc = b200_poseidon_merkle_dev(ctx, d.u(), cols, log_L, cap_height, dg.u());
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(cap_out, dg.u() + nd - (size_t(4) << cap_height), size_t(32) << cap_height, cudaMemcpyDeviceToHost, ctx->stream));
    if (digests_out) CUDA_TRY(cudaMemcpyAsync(digests_out, dg.p, nd * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
    API_END(ctx)
}

int b200_poseidon_permute(b200_ctx *ctx, uint64_t *states, size_t count) {
    API_BEGIN(ctx)
    if (!states || !count) throw ApiErr{B200_ERR_BAD_ARG, "b200_poseidon_permute: bad arguments"};
    DevBuf d(count * 96);
    CUDA_TRY(cudaMemcpyAsync(d.p, states, count * 96, cudaMemcpyHostToDevice, ctx->stream));
    poseidon_permute_batch(d.u(), count, ctx->stream);
    CUDA_TRY(cudaMemcpyAsync(states, d.p, count * 96, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
    API_END(ctx)
}

}  // extern "C"
