// Witness generation, first slice: the SimpleGenerators of the u32 custom gates, batched on the
// device.  In the reference every instance of these gates owns a generator object that the (sequential, CPU) witness
// generation loop runs once its input wires are known:
//   crypto/plonky2_u32/src/gates/arithmetic_u32.rs:379-430   U32ArithmeticGenerator      (m0, m1, addend -> lo, hi, inverse, 32 limbs)
//   crypto/plonky2_u32/src/gates/add_many_u32.rs:329-380     U32AddManyGenerator         (addends, carry_in -> result, carry_out, 16 + 2 limbs)
//   crypto/plonky2_u32/src/gates/subtraction_u32.rs:300-350  U32SubtractionGenerator     (x, y, borrow_in -> result, borrow_out, 16 limbs)
//   crypto/plonky2_u32/src/gates/range_check_u32.rs:200-240  U32RangeCheckGenerator      (input limbs -> 16 aux limbs each)
//   crypto/plonky2_u32/src/gates/comparison.rs:420-560       ComparisonGenerator         (first, second -> result, chunks, equality flags / dummies, msd, bits)
//   crypto/plonky2_u32/src/gates/interleave_u32.rs:289-318   U32InterleaveGenerator      (x -> 32 bits, x_interleaved)
//   crypto/plonky2_u32/src/gates/uninterleave_to_u32.rs:332-369, uninterleave_to_b32.rs  (x_interleaved -> 64 bits, evens, odds)
// A signature circuit has ~10^5 such instances and each is independent given its inputs, so one thread runs one
// (row, operation) generator over the column-major wire matrix in place.  The generators of the non-native / BigUint
// gadgets (plonky2_ed25519/src/gadgets/nonnative.rs:447-700, plonky2_ecdsa/src/gadgets/biguint.rs:417-470) and the
// dependency-ordered scheduling of generators stay on the Rust side of the boundary: the patched witness loop hands the
// rows of one dependency level to this call instead of running their generators one by one.
#include "../../include/b200_plonky2.h"
#include "ctx.h"
#include "host_common.h"

struct WitArgs { u64 *w; size_t n; const u32 *rows; u32 nrows; u32 p0, p1; };
#define WIRE(col) a.w[(size_t)(col) * a.n + row]

__global__ void __launch_bounds__(128) k_wit_u32_arithmetic(WitArgs a) {                // p0 = num_ops
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= (size_t)a.nrows * a.p0) return;
    const size_t row = a.rows[t / a.p0]; const u32 i = (u32)(t % a.p0), n = a.p0;
    const u64 out = gl_add(gl_mul(WIRE(6 * i), WIRE(6 * i + 1)), WIRE(6 * i + 2));     // canonical u64 of m0 * m1 + addend
    const u64 hi = out >> 32, lo = out & 0xFFFFFFFFULL;
    WIRE(6 * i + 3) = lo; WIRE(6 * i + 4) = hi;
    const u64 diff = 0xFFFFFFFFULL - hi;
    WIRE(6 * i + 5) = diff ? gl_inv(diff) : 0;
    for (int j = 0; j < 32; j++) WIRE(6 * n + 32 * i + j) = (out >> (2 * j)) & 3;
}
__global__ void __launch_bounds__(128) k_wit_u32_add_many(WitArgs a) {                  // p0 = num_addends, p1 = num_ops
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= (size_t)a.nrows * a.p1) return;
    const size_t row = a.rows[t / a.p1]; const u32 i = (u32)(t % a.p1), na = a.p0, n = a.p1, r = na + 3;
    u64 out = 0;
    for (u32 j = 0; j <= na; j++) out = gl_add(out, WIRE(r * i + j));                  // addends and carry_in (field sum, as the reference folds)
    const u64 res = out & 0xFFFFFFFFULL, car = out >> 32;
    WIRE(r * i + na + 1) = res; WIRE(r * i + na + 2) = car;
    for (int j = 0; j < 16; j++) WIRE(r * n + 18 * i + j) = (res >> (2 * j)) & 3;
    for (int j = 0; j < 2; j++) WIRE(r * n + 18 * i + 16 + j) = (car >> (2 * j)) & 3;
}
__global__ void __launch_bounds__(128) k_wit_u32_subtraction(WitArgs a) {               // p0 = num_ops
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= (size_t)a.nrows * a.p0) return;
    const size_t row = a.rows[t / a.p0]; const u32 i = (u32)(t % a.p0), n = a.p0;
    const u64 r0 = gl_sub(gl_sub(WIRE(5 * i), WIRE(5 * i + 1)), WIRE(5 * i + 2));
    const u64 bo = r0 > (1ULL << 32) ? 1 : 0;
    const u64 res = gl_add(r0, bo ? (1ULL << 32) : 0);
    WIRE(5 * i + 3) = res; WIRE(5 * i + 4) = bo;
    for (int j = 0; j < 16; j++) WIRE(5 * n + 16 * i + j) = (res >> (2 * j)) & 3;
}
__global__ void __launch_bounds__(128) k_wit_u32_range_check(WitArgs a) {               // p0 = num_input_limbs
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= (size_t)a.nrows * a.p0) return;
    const size_t row = a.rows[t / a.p0]; const u32 i = (u32)(t % a.p0), n = a.p0;
    const u64 x = WIRE(i);
    for (int j = 0; j < 16; j++) WIRE(n + 16 * i + j) = (x >> (2 * j)) & 3;
}
__global__ void __launch_bounds__(128) k_wit_u32_interleave(WitArgs a) {                // p0 = num_ops
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= (size_t)a.nrows * a.p0) return;
    const size_t row = a.rows[t / a.p0]; const u32 i = (u32)(t % a.p0), n = a.p0;
    const u64 x = WIRE(2 * i);
    u64 xi = 0;
    for (int j = 0; j < 32; j++) { const u64 bit = (x >> (31 - j)) & 1; WIRE(2 * n + 32 * i + j) = bit; xi += bit << (2 * (31 - j)); }
    WIRE(2 * i + 1) = xi;
}
__global__ void __launch_bounds__(128) k_wit_uninterleave(WitArgs a) {                  // p0 = num_ops, p1 = 1 for the B32 variant
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= (size_t)a.nrows * a.p0) return;
    const size_t row = a.rows[t / a.p0]; const u32 i = (u32)(t % a.p0), n = a.p0;
    const u64 x = WIRE(3 * i);
    u64 ev = 0, od = 0;
    for (int j = 0; j < 32; j++) {
        const int shift = 2 * (31 - j);
        const u64 e = (x >> (shift + 1)) & 1, o = (x >> shift) & 1;
        WIRE(3 * n + 64 * i + 2 * j) = e; WIRE(3 * n + 64 * i + 2 * j + 1) = o;
        const u64 coeff = a.p1 ? (1ULL << (2 * (31 - j))) : (1ULL << (31 - j));
        ev += e * coeff; od += o * coeff;
    }
    WIRE(3 * i + 1) = ev; WIRE(3 * i + 2) = od;
}
__global__ void __launch_bounds__(128) k_wit_comparison(WitArgs a) {                    // p0 = num_bits, p1 = num_chunks (one op per row)
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= a.nrows) return;
    const size_t row = a.rows[t]; const u32 nb = a.p0, nc = a.p1, cb = (nb + nc - 1) / nc;
    const u64 f = WIRE(0), s = WIRE(1);
    WIRE(2) = f <= s ? 1 : 0;
    u64 msd = 0;
    for (u32 i = 0; i < nc; i++) {
        const u64 fc = (f >> (cb * i)) & ((1ULL << cb) - 1), sc = (s >> (cb * i)) & ((1ULL << cb) - 1);
        WIRE(4 + i) = fc; WIRE(4 + nc + i) = sc;
        WIRE(4 + 2 * nc + i) = fc == sc ? 1 : gl_inv(gl_sub(sc, fc));                  // equality dummy
        WIRE(4 + 3 * nc + i) = fc == sc ? 1 : 0;                                        // chunks_equal
        if (fc != sc) { msd = gl_sub(sc, fc); WIRE(4 + 4 * nc + i) = 0; } else WIRE(4 + 4 * nc + i) = msd;     // intermediate values
    }
    WIRE(3) = msd;
    const u64 v = gl_add(1ULL << cb, msd);                                             // (two_n + most_significant_diff).to_canonical_u64(), comparison.rs:489-490
    for (u32 i = 0; i <= cb; i++) WIRE(4 + 5 * nc + i) = (v >> i) & 1;
}

int ctx_fail(b200_ctx *ctx, int code, const std::string &msg);
#define API_BEGIN(ctx) if (!(ctx)) return B200_ERR_BAD_ARG; try { cudaSetDevice((ctx)->device);
#define API_END(ctx) } catch (const CudaErr &e) { \
        char buf[256]; snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d", (int)e.e, cudaGetErrorString(e.e), e.file, e.line); \
        cudaGetLastError(); \
        return ctx_fail(ctx, e.e == cudaErrorMemoryAllocation ? B200_ERR_OOM : B200_ERR_CUDA, buf); \
    } catch (const ApiErr &e) { return ctx_fail(ctx, e.code, e.msg); \
    } catch (const std::bad_alloc &) { return ctx_fail(ctx, B200_ERR_OOM, "host allocation failed"); \
    } catch (const std::exception &e) { return ctx_fail(ctx, B200_ERR_BAD_ARG, e.what()); }

extern "C" {

int b200_witness_fill_u32_dev(b200_ctx *ctx, const b200_common *c, const uint32_t *row_gate, uint64_t *d_wires, uint32_t *rows_filled) {
    API_BEGIN(ctx)
    if (!c || !row_gate || !d_wires) throw ApiErr{B200_ERR_BAD_ARG, "b200_witness_fill_u32: null argument"};
    if (c->degree_bits > 26) throw ApiErr{B200_ERR_BAD_ARG, "degree_bits out of range"};
    const size_t n = size_t(1) << c->degree_bits;
    std::vector<std::vector<u32>> rows(c->num_gates);
    for (size_t i = 0; i < n; i++) { if (row_gate[i] >= c->num_gates) throw ApiErr{B200_ERR_BAD_ARG, "row_gate entry out of range"}; rows[row_gate[i]].push_back((u32)i); }
    u32 total = 0;
    std::vector<DevBuf> keep;
    for (u32 gi = 0; gi < c->num_gates; gi++) {
        GateDesc g; std::vector<u64> extra;
        if (rows[gi].empty() || !c->gate_ids[gi] || !parse_gate(c->gate_ids[gi], g, extra)) continue;
        WitArgs a; a.w = d_wires; a.n = n; a.nrows = (u32)rows[gi].size(); a.p0 = g.p[0]; a.p1 = g.p[1];
        size_t threads = 0; int kind = -1;
        switch (g.kind) {
        case G_U32_ARITHMETIC: case G_U32_SUBTRACTION: case G_U32_RANGE_CHECK: case G_U32_INTERLEAVE: threads = (size_t)a.nrows * g.p[0]; kind = (int)g.kind; break;
        case G_UNINTERLEAVE_U32: case G_UNINTERLEAVE_B32: threads = (size_t)a.nrows * g.p[0]; a.p1 = g.kind == G_UNINTERLEAVE_B32; kind = (int)g.kind; break;
        case G_U32_ADD_MANY: threads = (size_t)a.nrows * g.p[1]; kind = (int)g.kind; break;
        case G_COMPARISON: if (g.p[0] > 63) continue; threads = a.nrows; kind = (int)g.kind; break;
        default: continue;
        }
        keep.emplace_back((size_t)a.nrows * 4);
        CUDA_TRY(cudaMemcpyAsync(keep.back().p, rows[gi].data(), (size_t)a.nrows * 4, cudaMemcpyHostToDevice, ctx->stream));
        a.rows = (const u32 *)keep.back().p;
        const unsigned blocks = (unsigned)((threads + 127) / 128);
        switch (kind) {
        case G_U32_ARITHMETIC: k_wit_u32_arithmetic<<<blocks, 128, 0, ctx->stream>>>(a); break;
        case G_U32_ADD_MANY: k_wit_u32_add_many<<<blocks, 128, 0, ctx->stream>>>(a); break;
        case G_U32_SUBTRACTION: k_wit_u32_subtraction<<<blocks, 128, 0, ctx->stream>>>(a); break;
        case G_U32_RANGE_CHECK: k_wit_u32_range_check<<<blocks, 128, 0, ctx->stream>>>(a); break;
        case G_U32_INTERLEAVE: k_wit_u32_interleave<<<blocks, 128, 0, ctx->stream>>>(a); break;
        case G_UNINTERLEAVE_U32: case G_UNINTERLEAVE_B32: k_wit_uninterleave<<<blocks, 128, 0, ctx->stream>>>(a); break;
        case G_COMPARISON: k_wit_comparison<<<blocks, 128, 0, ctx->stream>>>(a); break;
        }
        LAUNCHED(1);
        total += a.nrows;
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));               // the row lists die with this call
    if (rows_filled) *rows_filled = total;
    return B200_OK;
    API_END(ctx)
}

int b200_witness_fill_u32(b200_ctx *ctx, const b200_common *c, const uint32_t *row_gate, uint64_t *wires, uint32_t *rows_filled) {
    API_BEGIN(ctx)
    if (!c || !wires) throw ApiErr{B200_ERR_BAD_ARG, "b200_witness_fill_u32: null argument"};
    if (c->degree_bits > 26) throw ApiErr{B200_ERR_BAD_ARG, "degree_bits out of range"};
    const size_t bytes = ((size_t)c->num_wires << c->degree_bits) * 8;
    DevBuf d(bytes);
    CUDA_TRY(cudaMemcpyAsync(d.p, wires, bytes, cudaMemcpyHostToDevice, ctx->stream));
    int rc = b200_witness_fill_u32_dev(ctx, c, row_gate, d.u(), rows_filled);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(wires, d.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
    API_END(ctx)
}

}  // extern "C"
