// Prover kernels: permutation partial products / Z, quotient evaluation on the 8n coset, polynomial openings,
// FRI batching + folding, proof-of-work grinding.  (loops C and D of section 8.)
// Replaces upstream plonky2 plonk/prover.rs::{wires_permutation_partial_products_and_zs, compute_quotient_polys},
// plonk/vanishing_poly.rs::eval_vanishing_poly_base_batch, fri/oracle.rs::prove_openings, fri/prover.rs.
#include "plonk.cuh"

// =====================================================================================================
// 1. partial products and Z
// =====================================================================================================
// per row i and challenge c: q_k = prod_{j in chunk k} (w_j + beta*k_j*x + gamma) / (w_j + beta*sigma_j + gamma)
// chunkq layout: [c][k][i] ; rowprod: [c][i]
__global__ void __launch_bounds__(128) k_chunk_quotients(const u64 *__restrict__ wires, const u64 *__restrict__ sigmas, const u64 *__restrict__ kis,
                                                         const u64 *__restrict__ tw, int log_n, u32 R, u32 npp, u32 qdf, u32 nch,
                                                         u64 b0, u64 g0, u64 b1, u64 g1, u64 *__restrict__ chunkq, u64 *__restrict__ rowprod) {
    size_t n = size_t(1) << log_n;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    // x = w_n^i from the twiddle table: tw[h + j] = w_{2h}^j, 2h = n
    u64 x;
    { size_t h = n >> 1; x = (i < h) ? tw[h + i] : gl_neg(tw[h + (i - h)]); if (n == 1) x = 1; }
    for (u32 c = 0; c < nch; c++) {
        u64 beta = c ? b1 : b0, gamma = c ? g1 : g0;
        u64 nums[16], dens[16];
        for (u32 k = 0; k <= npp; k++) {
            u64 num = 1, den = 1;
            for (u32 j = k * qdf; j < (k + 1) * qdf && j < R; j++) {
                u64 wv = wires[(size_t)j * n + i];
                u64 t = gl_add(wv, gamma);
                num = gl_mul(num, gl_add(t, gl_mul(beta, gl_mul(kis[j], x))));
                den = gl_mul(den, gl_add(t, gl_mul(beta, sigmas[(size_t)j * n + i])));
            }
            nums[k] = num; dens[k] = den;
        }
        // batch inversion of dens[0..npp]
        u64 pref[16]; u64 acc = 1;
        for (u32 k = 0; k <= npp; k++) { pref[k] = acc; acc = gl_mul(acc, dens[k]); }
        u64 inv = gl_inv(acc);
        u64 rp = 1;
        for (int k = (int)npp; k >= 0; k--) {
            u64 dinv = gl_mul(inv, pref[k]);
            inv = gl_mul(inv, dens[k]);
            u64 q = gl_mul(nums[k], dinv);
            chunkq[((size_t)c * (npp + 1) + k) * n + i] = q;
            rp = gl_mul(rp, q);
        }
        rowprod[(size_t)c * n + i] = rp;
    }
}

// exclusive multiplicative scan over rows, 1024 elements per block
#define SCAN_T 256
#define SCAN_E 4
// `in` and `excl` may alias (the prover scans the row products in place): no __restrict__ on either
__global__ void __launch_bounds__(SCAN_T) k_scan_local(const u64 *in, u64 *excl, u64 *__restrict__ block_tot, size_t n) {
    __shared__ u64 sh[SCAN_T];
    size_t c = blockIdx.y;
    size_t base = (size_t)blockIdx.x * SCAN_T * SCAN_E + threadIdx.x * SCAN_E;
    u64 v[SCAN_E]; u64 loc = 1;
#pragma unroll
    for (int e = 0; e < SCAN_E; e++) { v[e] = (base + e < n) ? in[c * n + base + e] : 1; loc = gl_mul(loc, v[e]); }
    sh[threadIdx.x] = loc;
    __syncthreads();
    // Hillis-Steele inclusive scan over thread totals
    for (int off = 1; off < SCAN_T; off <<= 1) {
        u64 t = (threadIdx.x >= off) ? sh[threadIdx.x - off] : 1;
        __syncthreads();
        sh[threadIdx.x] = gl_mul(sh[threadIdx.x], t);
        __syncthreads();
    }
    u64 pre = threadIdx.x ? sh[threadIdx.x - 1] : 1;
#pragma unroll
    for (int e = 0; e < SCAN_E; e++) { if (base + e < n) excl[c * n + base + e] = pre; pre = gl_mul(pre, v[e]); }
    if (threadIdx.x == SCAN_T - 1) block_tot[c * gridDim.x + blockIdx.x] = sh[SCAN_T - 1];
}
__global__ void k_scan_blocks(u64 *block_tot, u32 nblocks, u64 *final_prod) {   // one thread per challenge (nblocks small)
    u32 c = blockIdx.x;
    if (threadIdx.x != 0) return;
    u64 acc = 1;
    for (u32 b = 0; b < nblocks; b++) { u64 t = block_tot[c * nblocks + b]; block_tot[c * nblocks + b] = acc; acc = gl_mul(acc, t); }
    final_prod[c] = acc;
}
// Z[i] = blockoff * excl[i]; pp_k[i] = Z[i] * prod_{j<=k} q_j[i]
__global__ void __launch_bounds__(128) k_finish_zpp(const u64 *__restrict__ excl, const u64 *__restrict__ block_off, const u64 *__restrict__ chunkq,
                                                    size_t n, u32 npp, u32 nch, u32 nblocks, u64 *__restrict__ zs_pp) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (u32 c = 0; c < nch; c++) {
        u64 z = gl_mul(excl[(size_t)c * n + i], block_off[c * nblocks + i / (SCAN_T * SCAN_E)]);
        zs_pp[(size_t)c * n + i] = z;
        u64 acc = z;
        for (u32 k = 0; k < npp; k++) {
            acc = gl_mul(acc, chunkq[((size_t)c * (npp + 1) + k) * n + i]);
            zs_pp[((size_t)nch + (size_t)c * npp + k) * n + i] = acc;
        }
    }
}

void launch_partial_products(const u64 *d_wires, const u64 *d_sigmas, const u64 *d_kis, int log_n, u32 R, u32 npp, u32 qdf,
                             u32 nch, const u64 *betas, const u64 *gammas, const u64 *tw_fwd, u64 *d_zs_pp, u64 *d_scratch,
                             u64 *d_final_prod, cudaStream_t st) {
    size_t n = size_t(1) << log_n;
    u64 *chunkq = d_scratch;                               // nch*(npp+1)*n
    u64 *rowprod = chunkq + (size_t)nch * (npp + 1) * n;   // nch*n  (reused as excl)
    u32 nblocks = (u32)((n + SCAN_T * SCAN_E - 1) / (SCAN_T * SCAN_E));
    u64 *block_tot = rowprod + (size_t)nch * n;            // nch*nblocks
    k_chunk_quotients<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_wires, d_sigmas, d_kis, tw_fwd, log_n, R, npp, qdf, nch,
                                                                  betas[0], gammas[0], nch > 1 ? betas[1] : 0, nch > 1 ? gammas[1] : 0, chunkq, rowprod);
    LAUNCHED(1);
    k_scan_local<<<dim3(nblocks, nch), SCAN_T, 0, st>>>(rowprod, rowprod, block_tot, n);
    LAUNCHED(1);
    k_scan_blocks<<<nch, 32, 0, st>>>(block_tot, nblocks, d_final_prod);
    LAUNCHED(1);
    k_finish_zpp<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(rowprod, block_tot, chunkq, n, npp, nch, nblocks, d_zs_pp);
    LAUNCHED(1);
}

// =====================================================================================================
// 2. quotient
// =====================================================================================================
// --- 2a. permutation terms: acc[k][pos] = sum over the L_0 and partial-product checks of term * alpha_k^idx ---
__global__ void __launch_bounds__(128) k_quotient_perm(QuotientArgs a) {
    const int L = a.log_n + a.rate_bits;
    const size_t N = size_t(1) << L;
    size_t pos = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (pos >= N) return;
    const u32 i = bitrev32((u32)pos, L);                       // natural point index
    const u32 rmask = (1u << a.rate_bits) - 1;
    u64 x;
    { size_t h = N >> 1; u64 wi = (i < h) ? __ldg(a.tw_fwd + h + i) : gl_neg(__ldg(a.tw_fwd + h + (i - h))); x = gl_mul(GL_GEN, wi); }
    Wires cs = {a.cs_lde, N, pos}, w = {a.wires_lde, N, pos}, zs = {a.zs_lde, N, pos};
    const size_t pos_next = bitrev32((u32)((i + (1u << a.rate_bits)) & (N - 1)), L);
    Wires zs_next = {a.zs_lde, N, pos_next};
    const u32 nch = a.num_challenges, npp = a.npp;
    acc192 tot[MAX_CHALLENGES];
#pragma unroll
    for (int k = 0; k < MAX_CHALLENGES; k++) { tot[k].lo = 0; tot[k].hi = 0; tot[k].top = 0; }
    u64 l0 = gl_mul(a.zh_ev[i & rmask], gl_mul(a.n_inv, gl_inv(gl_sub(x, 1))));
    for (u32 c = 0; c < nch; c++) {
        u64 z = zs[c];
        u64 t = gl_mul(l0, gl_sub(z, 1));
        for (u32 k = 0; k < nch; k++) acc192_mad(tot[k], t, __ldg(a.apow + k * a.apow_stride + c));
        u64 prev = z;
        for (u32 kk = 0; kk <= npp; kk++) {
            u64 num = 1, den = 1;
            const u32 j0 = kk * a.qdf;
            if (a.qdf == 8 && j0 + 8 <= a.num_routed) {
                // the usual chunk of 8 routed wires: 16 independent column loads in flight before the product chains start
                u64 wv[8], sg[8];
                w.load<8>(j0, wv); cs.load<8>(a.num_constants + j0, sg);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    // products stay non-canonical between multiplications (any 64-bit representative is a valid factor)
                    u64 t0 = gl_add(wv[j], a.gammas[c]);
                    num = gl_mul_nc(num, gl_add_nc1(gl_mul_nc(__ldg(a.beta_kis + c * a.num_routed + j0 + j), x), t0));      // beta_c * k_j: one table per proof
                    den = gl_mul_nc(den, gl_add_nc1(gl_mul_nc(a.betas[c], sg[j]), t0));
                }
            } else {
                for (u32 j = j0; j < j0 + a.qdf && j < a.num_routed; j++) {
                    u64 t0 = gl_add(w[j], a.gammas[c]);
                    num = gl_mul(num, gl_add(t0, gl_mul(__ldg(a.beta_kis + c * a.num_routed + j), x)));
                    den = gl_mul(den, gl_add(t0, gl_mul(a.betas[c], cs[a.num_constants + j])));
                }
            }
            u64 next = kk < npp ? zs[nch + c * npp + kk] : zs_next[c];
            u64 term = gl_sub(gl_mul(prev, num), gl_mul(next, den));
            u32 tidx = nch + c * (npp + 1) + kk;
            for (u32 k = 0; k < nch; k++) acc192_mad(tot[k], term, __ldg(a.apow + k * a.apow_stride + tidx));
            prev = next;
        }
    }
    for (u32 k = 0; k < nch; k++) a.acc[(size_t)k * N + pos] = acc192_reduce(tot[k]);
}

// --- 2b. one kernel per gate kind: acc[k][pos] += alpha_k^{gate_off} * filter_g(x) * sum_i c_i alpha_k^i ---
// sub_shift = s > 0: the gate's own constraints have degree <= 2^(rate_bits - s) in the columns, so P_k(x) = alpha_k^{gate_off} *
// sum_i c_i(x) alpha_k^i is a polynomial of degree < N >> s and its values on the sub-coset g * H_{N >> s} determine it.  In the
// bit-reversed storage order that sub-coset is the FIRST N >> s positions: the kernel runs there only and stores P_k
// unfiltered in sub_out[k][pos]; launch_quotient extends P_k to the whole coset with two NTTs and k_gate_combine applies
// the (degree ~ group size) selector filters.  Exact in F_p, so the quotient values do not change.
template <int KIND>
__global__ void __launch_bounds__(128) k_gate(QuotientArgs a, GateDesc g, u32 sub_shift, u64 *__restrict__ sub_out) {
    __shared__ u64 sm[KIND == G_POSEIDON ? 12 * 128 : 1];
    const int L = a.log_n + a.rate_bits;
    const size_t N = size_t(1) << L, NP = N >> sub_shift;
    size_t pos = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const bool active = pos < NP;
    if (!active) pos = NP - 1;
    Wires cs = {a.cs_lde, N, pos}, w = {a.wires_lde, N, pos};
    Wires gc = {a.cs_lde + (size_t)a.num_selectors * N, N, pos};
    u64 filt = 1;
    if (!sub_out) {
        const u64 sel = cs[g.selector_index];
        for (u32 r = g.group_start; r < g.group_end; r++) if (r != g.row) filt = gl_mul(filt, gl_sub((u64)r, sel));
        if (a.num_selectors > 1) filt = gl_mul(filt, gl_sub(0xFFFFFFFFULL, sel));
    }
    ConstraintSink sink; sink.apow = a.apow; sink.apow_stride = a.apow_stride;
    sink.reset();
    eval_gate<KIND>(g, gc, w, a.pi_hash, a.extra, sink, &sm[KIND == G_POSEIDON ? threadIdx.x : 0], 128);
    if (!active) return;
    const u32 gate_off = a.num_challenges + a.num_challenges * (a.npp + 1);
    for (u32 k = 0; k < a.num_challenges; k++) {
        u64 v = gl_mul(sink.result(k), __ldg(a.apow + k * a.apow_stride + gate_off));
        if (sub_out) { sub_out[(size_t)k * NP + pos] = v; continue; }
        u64 *p = a.acc + (size_t)k * N + pos;
        *p = gl_add(*p, gl_mul(filt, v));
    }
}

// acc[k][pos] += sum over the sub-sampled gates of filter_g(pos) * ext_g[k][pos]
#define MAX_SUB_GATES 24
// ext: P_k on the whole coset [nch][N]; or (half-domain gates) lo = the evaluated even half, ext = the extended odd half, both [nch][N/2]
struct SubGates { u32 count; struct { u32 selector_index, group_start, group_end, row; const u64 *lo, *ext; } g[MAX_SUB_GATES]; };
__global__ void __launch_bounds__(256) k_gate_combine(QuotientArgs a, SubGates sg) {
    const size_t N = size_t(1) << (a.log_n + a.rate_bits);
    const size_t pos = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (pos >= N) return;
    acc192 tot[MAX_CHALLENGES];
#pragma unroll
    for (int k = 0; k < MAX_CHALLENGES; k++) { tot[k].lo = 0; tot[k].hi = 0; tot[k].top = 0; }
    for (u32 i = 0; i < sg.count; i++) {
        const u64 sel = __ldg(a.cs_lde + (size_t)sg.g[i].selector_index * N + pos);
        u64 filt = 1;
        for (u32 r = sg.g[i].group_start; r < sg.g[i].group_end; r++) if (r != sg.g[i].row) filt = gl_mul(filt, gl_sub((u64)r, sel));
        if (a.num_selectors > 1) filt = gl_mul(filt, gl_sub(0xFFFFFFFFULL, sel));
        const u64 *src; size_t stride = N, off = pos;
        if (sg.g[i].lo) { stride = N >> 1; src = pos < stride ? sg.g[i].lo : sg.g[i].ext; off = pos & (stride - 1); }
        else src = sg.g[i].ext;
        for (u32 k = 0; k < a.num_challenges; k++) acc192_mad(tot[k], filt, __ldg(src + (size_t)k * stride + off));
    }
    for (u32 k = 0; k < a.num_challenges; k++) { u64 *p = a.acc + (size_t)k * N + pos; *p = gl_add(*p, acc192_reduce(tot[k])); }
}

// --- 2c. out[k][natural i] = acc[k][pos] / Z_H(x) ---
__global__ void k_quotient_finish(QuotientArgs a) {
    const int L = a.log_n + a.rate_bits;
    const size_t N = size_t(1) << L;
    size_t pos = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (pos >= N) return;
    const u32 i = bitrev32((u32)pos, L);
    const u32 rmask = (1u << a.rate_bits) - 1;
    for (u32 k = 0; k < a.num_challenges; k++) a.out[(size_t)k * N + i] = gl_mul(a.acc[(size_t)k * N + pos], a.zh_inv[i & rmask]);
}

template <int KIND> static void launch_gate(const QuotientArgs &a, const GateDesc &g, u32 sub_shift, u64 *sub_out, cudaStream_t st) {
    const size_t NP = (size_t(1) << (a.log_n + a.rate_bits)) >> sub_shift;
    k_gate<KIND><<<(unsigned)((NP + 127) / 128), 128, 0, st>>>(a, g, sub_shift, sub_out);
    LAUNCHED(1);
}

// How far a gate's evaluation domain can shrink: the largest s with degree <= 2^(rate_bits - s), for gates heavy enough
// (by constraint count) that half or more of their evaluation time outweighs the two extension NTTs; 0 = evaluate everywhere.
u32 quotient_gate_sub_shift(const GateDesc &g, int log_n, int rate_bits) {
    if (g.num_constraints < 32 || g.kind == G_POSEIDON || log_n < 6) return 0;
    const u32 d = gate_degree(g);
    u32 s = 0;
    while ((int)s < rate_bits && s < 3 && (d << (s + 1)) <= (1u << rate_bits)) s++;
    return s;
}
size_t quotient_sub_scratch_u64(const GateDesc *gates, u32 num_gates, int log_n, int rate_bits, u32 nch) {
    const size_t N = size_t(1) << (log_n + rate_bits);
    size_t tot = 0; u32 cnt = 0;
    for (u32 i = 0; i < num_gates; i++) {
        const u32 s = quotient_gate_sub_shift(gates[i], log_n, rate_bits);
        if (!s || gates[i].num_constraints == 0 || cnt == MAX_SUB_GATES) continue;
        cnt++; tot += (size_t)nch * (2 * (N >> s) + (s == 1 ? N >> 1 : N)) + 96;
    }
    return tot;
}

void launch_quotient(const QuotientArgs &a, const GateDesc *h_gates, const NttTables &tw, const u64 *const unit_scale[4], u64 *d_sub_scratch, cudaStream_t st) {
    const int L = a.log_n + a.rate_bits;
    size_t N = size_t(1) << L;
    unsigned blocks = (unsigned)((N + 127) / 128);
    k_quotient_perm<<<blocks, 128, 0, st>>>(a);
    LAUNCHED(1);
    // sub-sampled gates of one shift s share three buffers: vals [cnt*nch][N>>s], coeffs (same), ext [cnt*nch][N]
    struct Class { u32 cnt = 0; u64 *vals = nullptr, *coeffs = nullptr, *ext = nullptr; } cls[4];
    u32 shift_of[256]; u32 slot_of[256];
    u32 n_sub = 0;
    for (u32 gi = 0; gi < a.num_gates && gi < 256; gi++) {
        shift_of[gi] = 0;
        if (h_gates[gi].num_constraints == 0 || !d_sub_scratch || n_sub == MAX_SUB_GATES) continue;
        const u32 s = quotient_gate_sub_shift(h_gates[gi], a.log_n, a.rate_bits);
        if (!s || !unit_scale[s]) continue;
        shift_of[gi] = s; slot_of[gi] = cls[s].cnt++; n_sub++;
    }
    {
        u64 *p = d_sub_scratch;
        for (u32 s = 1; s < 4; s++) if (cls[s].cnt) {
            const size_t rows = (size_t)cls[s].cnt * a.num_challenges;
            cls[s].vals = p; p += rows * (N >> s) + 32; cls[s].coeffs = p; p += rows * (N >> s) + 32; cls[s].ext = p; p += rows * (s == 1 ? N >> 1 : N) + 32;
        }
    }
    SubGates sg; sg.count = 0;
    for (u32 gi = 0; gi < a.num_gates; gi++) {
        const GateDesc &g = h_gates[gi];
        if (g.num_constraints == 0) continue;
        const u32 s = gi < 256 ? shift_of[gi] : 0;
        u64 *sub_out = nullptr;
        if (s) {
            sub_out = cls[s].vals + (size_t)slot_of[gi] * a.num_challenges * (N >> s);
            auto &e = sg.g[sg.count++];
            e.selector_index = g.selector_index; e.group_start = g.group_start; e.group_end = g.group_end; e.row = g.row;
            e.lo = s == 1 ? sub_out : nullptr;
            e.ext = cls[s].ext + (size_t)slot_of[gi] * a.num_challenges * (s == 1 ? N >> 1 : N);
        }
        switch (g.kind) {
#define GK(K) case K: launch_gate<K>(a, g, s, sub_out, st); break;
        GK(G_CONSTANT) GK(G_PUBLIC_INPUT) GK(G_ARITHMETIC) GK(G_ARITHMETIC_EXT) GK(G_MUL_EXT) GK(G_BASE_SUM) GK(G_REDUCING)
        GK(G_REDUCING_EXT) GK(G_RANDOM_ACCESS) GK(G_POSEIDON_MDS) GK(G_POSEIDON) GK(G_COSET_INTERPOLATION) GK(G_U32_ARITHMETIC)
        GK(G_U32_ADD_MANY) GK(G_U32_SUBTRACTION) GK(G_U32_RANGE_CHECK) GK(G_U32_INTERLEAVE) GK(G_UNINTERLEAVE_U32)
        GK(G_UNINTERLEAVE_B32) GK(G_COMPARISON)
#undef GK
        default: throw CudaErr{cudaErrorInvalidValue, __FILE__, __LINE__};
        }
    }
    if (sg.count) {
        for (u32 s = 1; s < 4; s++) if (cls[s].cnt) {
            const int Lm = L - (int)s; const size_t M = size_t(1) << Lm, rows = (size_t)cls[s].cnt * a.num_challenges;
            // values on H_M (storage order) -> coefficients of P(g X) -> values on H_N (= the whole coset, storage order)
            ntt_dif(tw, cls[s].vals, M, true, cls[s].coeffs, M, Lm, rows, 0, nullptr, true, gl_inv((u64)M % GL_P), st);
            // (half-domain gates: only the odd coset w_N * H_M is new -- one transform with block 1 of the scale table, (w_N)^m)
            if (s == 1) ntt_dif(tw, cls[s].coeffs, M, true, cls[s].ext, M, Lm, rows, 0, unit_scale[s] + M, false, 1, st);
            else ntt_dif(tw, cls[s].coeffs, M, true, cls[s].ext, N, Lm, rows << s, (int)s, unit_scale[s], false, 1, st);
        }
        k_gate_combine<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(a, sg);
        LAUNCHED(1);
    }
    k_quotient_finish<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(a);
    LAUNCHED(1);
}

// big[ci] holds (br order) the coefficients of Q(g X): coefficient k of Q = big-coeff_k * g^{-k} (1/N already applied).
// chunk m coefficient k' (k = m n + k') lives at position 8*br_n(k') + br_3(m).
// g^{-k} for k < n comes from two small tables (k = hi * 2^lo_bits + lo): a 64-step exponentiation per element made this
// copy kernel cost 0.4 ms per proof
__global__ void k_split_pow_tables(u64 *__restrict__ t_lo, u64 *__restrict__ t_hi, int lo_bits, int hi_bits) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    const u64 ginv = gl_inv(GL_GEN);
    if (i < (1u << lo_bits)) t_lo[i] = gl_pow(ginv, i);
    if (i < (1u << hi_bits)) t_hi[i] = gl_pow(ginv, (u64)i << lo_bits);
}
struct SplitPows { u64 gm[16]; };           // g^{-m n}, m < 2^rate_bits
__global__ void k_quotient_split(const u64 *__restrict__ big, u64 *__restrict__ chunks, int log_n, int rate_bits, u32 nch,
                                 const u64 *__restrict__ t_lo, const u64 *__restrict__ t_hi, int lo_bits, SplitPows sp) {
    const size_t n = size_t(1) << log_n, N = n << rate_bits;
    size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= (size_t)nch * N) return;
    u32 ci = (u32)(idx / N); size_t rest = idx % N;
    u32 m = (u32)(rest >> log_n), q = (u32)(rest & (n - 1));         // output: chunk m, br position q
    u32 kp = bitrev32(q, log_n);                                     // coefficient k = m n + kp
    u64 v = big[(size_t)ci * N + ((size_t)q << rate_bits) + bitrev32(m, rate_bits)];
    const u64 gk = gl_mul(sp.gm[m], gl_mul(__ldg(t_lo + (kp & ((1u << lo_bits) - 1))), __ldg(t_hi + (kp >> lo_bits))));
    chunks[((size_t)ci * (1u << rate_bits) + m) * n + q] = gl_mul(v, gk);
}
size_t quotient_split_scratch_u64(int log_n) { return (size_t(1) << ((log_n + 1) / 2)) + (size_t(1) << (log_n / 2)) + 16; }
void launch_quotient_split(const u64 *d_big, u64 *d_chunks, int log_n, int rate_bits, u32 nch, u64 *d_scratch, cudaStream_t st) {
    const int lo_bits = (log_n + 1) / 2, hi_bits = log_n - lo_bits;
    u64 *t_lo = d_scratch, *t_hi = d_scratch + (size_t(1) << lo_bits);
    k_split_pow_tables<<<(unsigned)(((1u << lo_bits) + 127) / 128), 128, 0, st>>>(t_lo, t_hi, lo_bits, hi_bits);
    LAUNCHED(1);
    SplitPows sp; const u64 gn = gl_pow(gl_inv(GL_GEN), u64(1) << log_n);
    { u64 x = 1; for (int m = 0; m < 16; m++) { sp.gm[m] = x; x = gl_mul(x, gn); } }
    size_t total = (size_t)nch << (log_n + rate_bits);
    k_quotient_split<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_big, d_chunks, log_n, rate_bits, nch, t_lo, t_hi, lo_bits, sp);
    LAUNCHED(1);
}

// =====================================================================================================
// 3. openings
// =====================================================================================================
struct Pows2 { gl2 v[32]; };
__global__ void k_ext_pow_table(gl2 *table, int L, Pows2 p2) {
    size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (q >= (size_t(1) << L)) return;
    u32 e = bitrev32((u32)q, L);
    gl2 r = gl2_make(1, 0);
    for (int b = 0; b < L; b++) if (e & (1u << b)) r = gl2_mul(r, p2.v[b]);
    table[q] = r;
}
void launch_ext_pow_table(gl2 *d_table, int L, const gl2 *h_pows2, cudaStream_t st) {
    Pows2 p; for (int b = 0; b < 32; b++) p.v[b] = b < L ? h_pows2[b] : gl2_make(0, 0);
    size_t n = size_t(1) << L;
    k_ext_pow_table<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_table, L, p);
    LAUNCHED(1);
}
// out[p] = sum_q coeffs[p][q] * table[q].  The table (powers of the opening point) is the same for every polynomial: a block takes
// EP_POLYS polynomials and one of EP_SEGS segments of the index range, so that a table entry is loaded once per EP_POLYS
// products and the grid has polys/EP_POLYS x EP_SEGS blocks (one block per polynomial left two thirds of the SMs idle at 355
// polynomials); a second tiny kernel adds the EP_SEGS partial sums.
#define EP_POLYS 4
#define EP_SEGS 16
__global__ void __launch_bounds__(256) k_eval_polys(const u64 *__restrict__ coeffs, u32 num_polys, int L, const gl2 *__restrict__ table, gl2 *__restrict__ partial) {
    __shared__ u64 sa[256], sb[256];
    const size_t n = size_t(1) << L;
    const u32 p0 = blockIdx.x * EP_POLYS, seg = blockIdx.y;
    const size_t seg_len = (n + EP_SEGS - 1) / EP_SEGS, q0 = seg * seg_len, q1 = q0 + seg_len < n ? q0 + seg_len : n;
    acc192 A[EP_POLYS], B[EP_POLYS];
#pragma unroll
    for (int k = 0; k < EP_POLYS; k++) { A[k].lo = A[k].hi = 0; A[k].top = 0; B[k].lo = B[k].hi = 0; B[k].top = 0; }
    for (size_t q = q0 + threadIdx.x; q < q1; q += 256) {
        const gl2 t = table[q];
        u64 v[EP_POLYS];
#pragma unroll
        for (int k = 0; k < EP_POLYS; k++) v[k] = (p0 + k < num_polys) ? coeffs[(size_t)(p0 + k) * n + q] : 0;
#pragma unroll
        for (int k = 0; k < EP_POLYS; k++) { acc192_mad(A[k], v[k], t.a); acc192_mad(B[k], v[k], t.b); }
    }
#pragma unroll
    for (int k = 0; k < EP_POLYS; k++) {
        sa[threadIdx.x] = acc192_reduce(A[k]); sb[threadIdx.x] = acc192_reduce(B[k]);
        __syncthreads();
        for (int off = 128; off > 0; off >>= 1) {
            if (threadIdx.x < off) { sa[threadIdx.x] = gl_add(sa[threadIdx.x], sa[threadIdx.x + off]); sb[threadIdx.x] = gl_add(sb[threadIdx.x], sb[threadIdx.x + off]); }
            __syncthreads();
        }
        if (threadIdx.x == 0 && p0 + k < num_polys) partial[(size_t)(p0 + k) * EP_SEGS + seg] = gl2_make(sa[0], sb[0]);
        __syncthreads();
    }
}
__global__ void k_eval_polys_sum(const gl2 *__restrict__ partial, u32 num_polys, gl2 *__restrict__ out) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= num_polys) return;
    gl2 s = gl2_make(0, 0);
    for (int k = 0; k < EP_SEGS; k++) s = gl2_add(s, partial[(size_t)p * EP_SEGS + k]);
    out[p] = s;
}
void launch_eval_polys(const u64 *d_coeffs_br, size_t num_polys, int L, const gl2 *d_table, gl2 *d_out, gl2 *d_partial, cudaStream_t st) {
    k_eval_polys<<<dim3((unsigned)((num_polys + EP_POLYS - 1) / EP_POLYS), EP_SEGS), 256, 0, st>>>(d_coeffs_br, (u32)num_polys, L, d_table, d_partial);
    LAUNCHED(1);
    k_eval_polys_sum<<<(unsigned)((num_polys + 127) / 128), 128, 0, st>>>(d_partial, (u32)num_polys, d_out);
    LAUNCHED(1);
}
size_t eval_polys_scratch_gl2(size_t num_polys) { return num_polys * EP_SEGS; }

// =====================================================================================================
// 4. FRI: batching of all oracles into one extension-field codeword, folding, proof of work
// =====================================================================================================
// Step 1, coefficient space: G0 = sum_p alpha^p f_p, G1 = sum_{c < nch} alpha^c Z_c as coefficient vectors (the same
// bit-reversed coefficient order as every committed polynomial): g[0..3][q] = G0.a, G0.b, G1.a, G1.b.  The combination is
// F-linear, so combining the n coefficients and extending the 4 resulting columns equals combining the 8n LDE values of all
// 355 columns (what round 1 did: 3 GB of reads per proof instead of 0.37 GB).
__global__ void __launch_bounds__(128) k_fri_coeff_combine(FriCombineArgs a) {
    const size_t n = size_t(1) << a.log_n;
    const size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (q >= n) return;
    acc192 A = {0, 0, 0}, B = {0, 0, 0};
    u32 p = 0;
    for (int o = 0; o < 4; o++) {
        const u64 *base = a.coeffs[o] + q;
#pragma unroll 8
        for (u32 c = 0; c < a.cols[o]; c++, p++) {
            u64 v = __ldg(base + (size_t)c * n);
            gl2 ap = a.apow[p];
            acc192_mad(A, v, ap.a); acc192_mad(B, v, ap.b);
        }
    }
    acc192 A1 = {0, 0, 0}, B1 = {0, 0, 0};
    for (u32 c = 0; c < a.nch; c++) {
        u64 v = __ldg(a.coeffs[2] + (size_t)c * n + q);
        gl2 ap = a.apow[c];
        acc192_mad(A1, v, ap.a); acc192_mad(B1, v, ap.b);
    }
    a.g[q] = acc192_reduce(A); a.g[n + q] = acc192_reduce(B); a.g[2 * n + q] = acc192_reduce(A1); a.g[3 * n + q] = acc192_reduce(B1);
}
// Step 3, on the LDE of g: out(x) = alpha^nch * (G0(x) - y0) / (x - zeta) + (G1(x) - y1) / (x - g zeta)
__global__ void __launch_bounds__(128) k_fri_combine(FriCombineArgs a) {
    const size_t N = size_t(1) << a.log_N;
    size_t pos = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (pos >= N) return;
    const u32 i = bitrev32((u32)pos, a.log_N);
    u64 x;
    { size_t h = N >> 1; u64 wi = (i < h) ? __ldg(a.tw_fwd + h + i) : gl_neg(__ldg(a.tw_fwd + h + (i - h))); x = gl_mul(GL_GEN, wi); }
    const gl2 F0 = gl2_make(__ldg(a.g_lde + pos), __ldg(a.g_lde + N + pos));
    const gl2 F1 = gl2_make(__ldg(a.g_lde + 2 * N + pos), __ldg(a.g_lde + 3 * N + pos));
    const gl2 xe = gl2_make(x, 0);
    const gl2 d0 = gl2_sub(xe, a.zeta), d1 = gl2_sub(xe, a.gzeta);
    const gl2 mi = gl2_inv(gl2_mul(d0, d1));                    // one inversion for both denominators
    gl2 t0 = gl2_mul(gl2_sub(F0, a.y0), gl2_mul(mi, d1));
    gl2 t1 = gl2_mul(gl2_sub(F1, a.y1), gl2_mul(mi, d0));
    a.out[pos] = gl2_add(gl2_mul(t0, a.alpha_nch), t1);
}
void launch_fri_coeff_combine(const FriCombineArgs &a, cudaStream_t st) {
    size_t n = size_t(1) << a.log_n;
    k_fri_coeff_combine<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(a);
    LAUNCHED(1);
}
void launch_fri_combine(const FriCombineArgs &a, cudaStream_t st) {
    size_t N = size_t(1) << a.log_N;
    k_fri_combine<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(a);
    LAUNCHED(1);
}

// one thread per coset of 2^ab values (consecutive in bit-reversed order)
template <int AB>
__global__ void __launch_bounds__(128) k_fri_fold(const gl2 *__restrict__ in, gl2 *__restrict__ out, int log_len, gl2 beta, u64 shift,
                                                  const u64 *__restrict__ tw) {
    constexpr int A = 1 << AB;
    const size_t len = size_t(1) << log_len, nout = len >> AB;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= nout) return;
    // values in natural coset order: y[m] = in[A*i + br_AB(m)], at points x0 * w_A^m, x0 = shift * w_len^{br(i)}
    gl2 y[A];
#pragma unroll
    for (int k = 0; k < A; k++) y[bitrev32(k, AB)] = in[A * i + k];
    // inverse DFT of size A (DIF with inverse twiddles, output bit-reversed), unnormalised
    u64 winv[A / 2];
    { u64 wa = gl_inv(gl_root_of_unity(AB)); winv[0] = 1; for (int k = 1; k < A / 2; k++) winv[k] = gl_mul(winv[k - 1], wa); }
#pragma unroll
    for (int s = AB - 1; s >= 0; s--) {
        const int h = 1 << s;
#pragma unroll
        for (int e = 0; e < A; e++) {
            if (e & h) continue;
            int j = e & (h - 1);
            gl2 u = y[e], v = y[e + h];
            y[e] = gl2_add(u, v);
            y[e + h] = gl2_scale(gl2_sub(u, v), winv[j * (A / 2 / h)]);
        }
    }
    // d_t = y[br(t)] / A ; folded = sum_t (beta / x0)^t d_t
    u32 e = bitrev32((u32)i, log_len - AB);
    u64 wl;
    { size_t hh = len >> 1; wl = (len == 1) ? 1 : ((e < hh) ? __ldg(tw + hh + e) : gl_neg(__ldg(tw + hh + (e - hh)))); }
    u64 x0 = gl_mul(shift, wl);
    gl2 r = gl2_scale(beta, gl_inv(x0));
    gl2 acc = gl2_make(0, 0);
#pragma unroll
    for (int t = A - 1; t >= 0; t--) acc = gl2_add(gl2_mul(acc, r), y[bitrev32(t, AB)]);
    out[i] = gl2_scale(acc, gl_inv((u64)A));
}
void launch_fri_fold(const gl2 *d_in, gl2 *d_out, int log_len, int arity_bits, gl2 beta, u64 shift, const u64 *tw_fwd, cudaStream_t st) {
    size_t nout = size_t(1) << (log_len - arity_bits);
    unsigned blocks = (unsigned)((nout + 127) / 128);
    switch (arity_bits) {
    case 1: k_fri_fold<1><<<blocks, 128, 0, st>>>(d_in, d_out, log_len, beta, shift, tw_fwd); break;
    case 2: k_fri_fold<2><<<blocks, 128, 0, st>>>(d_in, d_out, log_len, beta, shift, tw_fwd); break;
    case 3: k_fri_fold<3><<<blocks, 128, 0, st>>>(d_in, d_out, log_len, beta, shift, tw_fwd); break;
    case 4: k_fri_fold<4><<<blocks, 128, 0, st>>>(d_in, d_out, log_len, beta, shift, tw_fwd); break;
    default: throw CudaErr{cudaErrorInvalidValue, __FILE__, __LINE__};
    }
    LAUNCHED(1);
}

struct PowState { u64 s[12]; };
__global__ void __launch_bounds__(128) k_pow(PowState base, int pos, u64 start, u64 count, int bits, unsigned long long *result) {
    __shared__ u64 sm[12 * 128];                    // state word i of thread t at sm[128 * i + t] (conflict-free, as in merkle.cu)
    u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    u64 cand = start + t;
    const bool live = t < count && cand < GL_P;
    u64 *st = sm + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 12; i++) st[128 * i] = (i == pos) ? cand : base.s[i];
    poseidon_permute_smem(st, 128);
    if (live && (st[128 * 7] >> (64 - bits)) == 0) atomicMin(result, (unsigned long long)cand);
}
void launch_pow(const u64 state[12], int pos, u64 start, u64 count, int bits, u64 *d_result, cudaStream_t st) {
    PowState b; for (int i = 0; i < 12; i++) b.s[i] = state[i];
    k_pow<<<(unsigned)((count + 127) / 128), 128, 0, st>>>(b, pos, start, count, bits, (unsigned long long *)d_result);
    LAUNCHED(1);
}
