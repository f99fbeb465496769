// b200_circuit_upload / b200_prove: the GPU plonky2 prover (PoseidonGoldilocksConfig, D = 2, zero_knowledge = false,
// no lookups).  Replaces upstream plonky2::plonk::prover::prove after witness generation
// (call sites: near_bft_finality/src/prove_crypto/ed25519.rs:60,100, recursion.rs:95, sha256.rs:82,
//  prove_block_data/keys_stakes.rs:244, primitives.rs:108,158,343) and the constants/sigmas commitment of
// CircuitBuilder::build (ed25519.rs:33, recursion.rs:94, sha256.rs:78).
//
// Pipeline (all device-resident, column-major, LDE rows bit-reversed; host only runs the Fiat-Shamir sponge):
//   wires --iNTT--> coeffs --coset LDE--> [W][8n] --Poseidon leaves/Merkle--> cap  ==> beta, gamma
//   partial products + Z (scan) --commit--> cap                                   ==> alpha
//   quotient on the 8n coset (gate table) --iNTT--> 8 chunks/challenge --commit--> cap ==> zeta
//   openings (dot products with zeta powers)                                      ==> fri alpha
//   FRI: value-space batching of all four oracles, arity-16 folds + Merkle per layer ==> betas
//   final poly (host, tiny), 16-bit PoW grind (min nonce), 28 query openings (gather kernel)
#include "../../include/b200_plonky2.h"
#include "ctx.h"
#include "plonk.cuh"
#include "poseidon.cuh"
#include "bn254.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

#include "host_common.h"
#include "vanishing.h"

// ------------------------------------------------------------------------------------------------
struct b200_circuit {
    b200_ctx *ctx;
    u32 num_wires, num_routed, num_constants, num_challenges, degree_bits, rate_bits, cap_height, pow_bits, num_queries,
        qdf, npp, num_gate_constraints, num_pi, num_selectors, hasher;
    std::vector<u32> arity_bits;
    std::vector<GateDesc> gates;
    std::vector<u64> k_is, extra;       // extra: per-gate constant tables (barycentric weights), host copy
    DevBuf d_gates, d_extra, d_kis;
    DevBuf cs_coeffs, cs_lde, cs_digests, sigma_values;
    u64 digest[4];
    std::vector<u64> cap;               // constants/sigmas cap, host copy
    // circuit cache (b200_circuit_build): handles are reference counted, the context keeps cached circuits alive
    int refs = 1; bool cached = false; u64 key[2] = {0, 0}; u64 last_use = 0;
};

struct Arena {      // bump allocator over one device allocation, reset per prove
    u8 *base = nullptr; size_t cap = 0, off = 0;
    u64 *take(size_t n_u64) { size_t b = (n_u64 * 8 + 255) & ~size_t(255); if (off + b > cap) throw ApiErr{B200_ERR_OOM, "arena exhausted"}; u64 *p = (u64 *)(base + off); off += b; return p; }
};

// Small host<->device transfers go through one pinned staging buffer per context: copies to/from pageable memory are
// synchronous and staged by the driver (tens of microseconds each, ~15 per proof), which matters for the small
// recursion-shaped proofs whose whole prove() takes a few milliseconds.
struct Staging {
    b200_ctx *ctx; cudaStream_t st; size_t off = 0;
    struct Pending { void *dst; const void *src; size_t bytes; };
    std::vector<Pending> pending;
    static constexpr size_t CAP = size_t(8) << 20;
    Staging(b200_ctx *c) : ctx(c), st(c->stream) {
        if (!ctx->pinned) CUDA_TRY(cudaMallocHost(&ctx->pinned, CAP));
    }
    u8 *take(size_t bytes) { size_t b = (bytes + 63) & ~size_t(63); if (off + b > CAP) { sync(); off = 0; } if (b > CAP) throw ApiErr{B200_ERR_OOM, "staging buffer too small"}; u8 *p = (u8 *)ctx->pinned + off; off += b; return p; }
    void d2h(void *dst, const void *dsrc, size_t bytes) {
        u8 *p = take(bytes);
        CUDA_TRY(cudaMemcpyAsync(p, dsrc, bytes, cudaMemcpyDeviceToHost, st));
        pending.push_back({dst, p, bytes});
    }
    void h2d(void *ddst, const void *src, size_t bytes) {
        u8 *p = take(bytes); memcpy(p, src, bytes);
        CUDA_TRY(cudaMemcpyAsync(ddst, p, bytes, cudaMemcpyHostToDevice, st));
    }
    void sync() {
        if (ctx->timing) {
            auto t0 = std::chrono::steady_clock::now();
            if (ctx->host_seg_ms.size() <= ctx->host_seg_i) ctx->host_seg_ms.resize(ctx->host_seg_i + 1, 0.0);
            ctx->host_seg_ms[ctx->host_seg_i++] += std::chrono::duration<double, std::milli>(t0 - ctx->host_seg_t0).count();
            CUDA_TRY(cudaStreamSynchronize(st));
            ctx->host_seg_t0 = std::chrono::steady_clock::now();
            ctx->host_wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); ctx->host_syncs++;
        } else CUDA_TRY(cudaStreamSynchronize(st));
        for (auto &q : pending) memcpy(q.dst, q.src, q.bytes);
        pending.clear();
    }
};

static void put_u64(std::vector<uint8_t> &o, u64 x) { for (int i = 0; i < 8; i++) o.push_back((uint8_t)(x >> (8 * i))); }
static void put_words(std::vector<uint8_t> &o, const u64 *w, size_t n) {
#if defined(__BYTE_ORDER__) && __BYTE_ORDER__ == __ORDER_LITTLE_ENDIAN__
    const uint8_t *b = (const uint8_t *)w; o.insert(o.end(), b, b + 8 * n);       // the proof format is little-endian u64 words
#else
    for (size_t i = 0; i < n; i++) put_u64(o, w[i]);
#endif
}

// host-side tiny NTT (final polynomial only)
static void host_ntt(std::vector<u64> &a, int logn, bool inverse) {
    size_t n = size_t(1) << logn;
    for (size_t i = 0; i < n; i++) { size_t j = bitrev32((u32)i, logn); if (i < j) std::swap(a[i], a[j]); }
    u64 wn = gl_root_of_unity(logn); if (inverse) wn = gl_inv(wn);
    for (int s = 1; s <= logn; s++) {
        size_t m = size_t(1) << s, h = m >> 1; u64 wm = gl_pow(wn, n >> s);
        for (size_t k = 0; k < n; k += m) { u64 w = 1; for (size_t j = 0; j < h; j++) { u64 t = gl_mul(w, a[k + j + h]), u = a[k + j]; a[k + j] = gl_add(u, t); a[k + j + h] = gl_sub(u, t); w = gl_mul(w, wm); } }
    }
    if (inverse) { u64 ni = gl_inv((u64)n % GL_P); for (auto &x : a) x = gl_mul(x, ni); }
}

struct CopyDesc { const u64 *src; u64 stride; u32 count; u32 dst_off; };
__global__ void k_copy_descs(const CopyDesc *__restrict__ d, u64 *__restrict__ dst) {
    CopyDesc c = d[blockIdx.x];
    for (u32 i = threadIdx.x; i < c.count; i += blockDim.x) dst[c.dst_off + i] = c.src[(size_t)i * c.stride];
}

static int prove_impl(b200_ctx *ctx, const b200_circuit *k, const u64 *wires_in, bool wires_on_device, const u64 *pis, size_t n_pis,
                      u64 pow_hint, uint8_t **proof_out, size_t *proof_len) {
    cudaStream_t st = ctx->stream;
    struct WallTimer { b200_ctx *c; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
                       ~WallTimer() { if (c->timing) { c->host_total_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); c->host_proofs++; } } } wall_timer{ctx};
    ctx->host_seg_i = 0; ctx->host_seg_t0 = std::chrono::steady_clock::now();
    const int nlog = (int)k->degree_bits, rb = (int)k->rate_bits, chh = (int)k->cap_height;
    const int Llog = nlog + rb;
    const size_t n = size_t(1) << nlog, N = n << rb;
    const u32 W = k->num_wires, R = k->num_routed, NC = k->num_constants, nch = k->num_challenges, qdf = k->qdf, npp = k->npp;
    const u32 nzp = nch * (1 + npp), nq = nch * qdf;
    const size_t capw = size_t(4) << chh, ncap = size_t(1) << chh;
    const int HS = (int)k->hasher;
    if (n_pis != k->num_pi) throw ApiErr{B200_ERR_BAD_ARG, "b200_prove: wrong number of public inputs"};
    ntt_tables_ensure(ctx->tw, Llog, st);

    // ---- arena layout ----
    const size_t dig_u64 = merkle_num_u64(Llog, chh);
    const size_t pp_scratch_u64 = partial_products_scratch_u64(nlog, npp, nch);
    size_t need = 0;
    auto rnd = [](size_t u) { return (u * 8 + 255) & ~size_t(255); };
    need += (wires_on_device ? 0 : rnd((size_t)W * n)) + rnd((size_t)W * n) + rnd((size_t)W * N) + rnd(dig_u64);      // wires
    need += 2 * rnd((size_t)nzp * n) + rnd((size_t)nzp * N) + rnd(dig_u64) + rnd(pp_scratch_u64) + rnd(8);   // zs
    need += rnd((size_t)nch * N) * 2 + rnd((size_t)nq * n) + rnd((size_t)nq * N) + rnd(dig_u64);                        // quotient
    const size_t sub_scratch_u64 = quotient_sub_scratch_u64(k->gates.data(), (u32)k->gates.size(), nlog, rb, nch);
    need += rnd(sub_scratch_u64) + rnd((size_t)nch * R + 8) + rnd(quotient_split_scratch_u64(nlog));
    need += rnd(2 * n) * 2 + rnd(2 * (size_t)(NC + R + W + nzp + nq + nch)) + rnd(2 * (size_t)(NC + R + W + nzp + nq)) + rnd(2 * eval_polys_scratch_gl2(std::max<size_t>(std::max<size_t>(W, NC + R), std::max<size_t>(nzp, nq)))); // openings tables
    need += rnd(2 * N) * 2 + rnd(dig_u64) * 2 + rnd(1 << 20) + rnd(4 * n) + rnd(4 * N);                                                            // fri
    Arena A; A.base = ctx_arena(ctx, need + (1 << 20)); A.cap = need + (1 << 20);
    Staging S(ctx);

    // ---- public input hash, wires commitment ----
    u64 pi_hash[4]; poseidon_hash_no_pad_host(pis, n_pis, pi_hash);
    const u64 *d_wires;
    if (wires_on_device) d_wires = wires_in;
    struct CopyDrain { cudaStream_t s = nullptr; ~CopyDrain() { if (s) cudaStreamSynchronize(s); } } drain;
    u64 *w_coeffs = A.take((size_t)W * n), *w_lde = A.take((size_t)W * N), *w_dig = A.take(dig_u64);
    if (wires_on_device) commit_values(ctx, d_wires = wires_in, W, nlog, rb, chh, w_coeffs, w_lde, w_dig, HS);
    else {
        // host witness: stream it in column chunks on a copy stream and run iNTT + LDE of chunk c while chunk c+1 is in
        // flight (the LDE takes longer than the PCIe copy, so the H2D time disappears behind it); hash once all columns are in
        u64 *t = A.take((size_t)W * n); d_wires = t;
        if (!ctx->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        const u32 nchunks = W >= 32 ? (u32)std::min<size_t>(6, std::max<size_t>(1, ((size_t)W * n) >> 19)) : 1;   // >= 2^19 words per chunk: small proofs are launch-bound
        // RAII: the events die and the copy stream is drained on every exit path (the caller's witness buffer must not
        // have an H2D copy in flight once b200_prove returns, error or not)
        struct Ev { cudaEvent_t e = nullptr; Ev() { CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); } ~Ev() { if (e) cudaEventDestroy(e); }
                    Ev(const Ev &) = delete; Ev &operator=(const Ev &) = delete; };
        drain.s = ctx->copy_stream;
        std::vector<std::unique_ptr<Ev>> evs;
        for (u32 c = 0; c < nchunks; c++) evs.emplace_back(new Ev());
        Ev start_ev;
        CUDA_TRY(cudaEventRecord(start_ev.e, st)); CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, start_ev.e, 0));   // arena reuse ordering
        for (u32 c = 0; c < nchunks; c++) {
            size_t c0 = (size_t)W * c / nchunks, c1 = (size_t)W * (c + 1) / nchunks;
            CUDA_TRY(cudaMemcpyAsync(t + c0 * n, wires_in + c0 * n, (c1 - c0) * n * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
            CUDA_TRY(cudaEventRecord(evs[c]->e, ctx->copy_stream));
        }
        for (u32 c = 0; c < nchunks; c++) {
            size_t c0 = (size_t)W * c / nchunks, c1 = (size_t)W * (c + 1) / nchunks;
            CUDA_TRY(cudaStreamWaitEvent(st, evs[c]->e, 0));
            commit_values_lde_only(ctx, t + c0 * n, c1 - c0, nlog, rb, w_coeffs + c0 * n, w_lde + c0 * N);
        }
        commit_hash_only(ctx, w_lde, W, nlog, rb, chh, w_dig, HS);
    }
    std::vector<u64> cap_w(capw), cap_z(capw), cap_q(capw);
    S.d2h(cap_w.data(), w_dig + dig_u64 - capw, capw * 8);
    S.sync();
    drain.s = nullptr;          // the compute stream waited on every chunk's event: the witness copies are complete

    Challenger ch;
    ch.observe_hash(k->digest, HS); ch.observe_n(pi_hash, 4); ch.observe_cap(cap_w.data(), ncap, HS);
    u64 betas[MAX_CHALLENGES] = {0, 0}, gammas[MAX_CHALLENGES] = {0, 0}, alphas[MAX_CHALLENGES] = {0, 0};
    for (u32 i = 0; i < nch; i++) betas[i] = ch.get();
    for (u32 i = 0; i < nch; i++) gammas[i] = ch.get();

    // ---- partial products + Z ----
    u64 *zs_vals = A.take((size_t)nzp * n), *zs_coeffs = A.take((size_t)nzp * n), *zs_lde = A.take((size_t)nzp * N), *zs_dig = A.take(dig_u64);
    u64 *pp_scratch = A.take(pp_scratch_u64), *d_final_prod = A.take(8);
    { PhaseTimer t(ctx, "partial_products");
    launch_partial_products(d_wires, k->sigma_values.u(), k->d_kis.u(), nlog, R, npp, qdf, nch, betas, gammas, ctx->tw.fwd, zs_vals, pp_scratch, d_final_prod, st); }
    commit_values(ctx, zs_vals, nzp, nlog, rb, chh, zs_coeffs, zs_lde, zs_dig, HS);
    u64 final_prod[MAX_CHALLENGES];
    S.d2h(cap_z.data(), zs_dig + dig_u64 - capw, capw * 8);
    S.d2h(final_prod, d_final_prod, nch * 8);
    S.sync();
    for (u32 i = 0; i < nch; i++) if (final_prod[i] != 1) throw ApiErr{B200_ERR_UNSATISFIED, "permutation argument: grand product != 1 (copy constraints violated)"};
    ch.observe_cap(cap_z.data(), ncap, HS);
    for (u32 i = 0; i < nch; i++) alphas[i] = ch.get();

    // ---- quotient ----
    const u32 gate_off = nch + nch * (npp + 1);
    const u32 apow_stride = gate_off + k->num_gate_constraints + 1;
    {
        std::vector<u64> apow((size_t)MAX_CHALLENGES * apow_stride, 0);
        for (u32 c = 0; c < nch; c++) { u64 p = 1; for (u32 i = 0; i < apow_stride; i++) { apow[(size_t)c * apow_stride + i] = p; p = gl_mul(p, alphas[c]); } }
        u64 *d_apow = A.take(apow.size());
        S.h2d(d_apow, apow.data(), apow.size() * 8);
        u64 *q_vals = A.take((size_t)nch * N), *q_big = A.take((size_t)nch * N);
        QuotientArgs qa; memset(&qa, 0, sizeof qa);
        qa.cs_lde = k->cs_lde.u(); qa.wires_lde = w_lde; qa.zs_lde = zs_lde; qa.out = q_vals;
        qa.gates = (const GateDesc *)k->d_gates.p; qa.extra = k->d_extra.u(); qa.k_is = k->d_kis.u(); qa.apow = d_apow; qa.tw_fwd = ctx->tw.fwd;
        qa.apow_stride = apow_stride; qa.num_gates = (u32)k->gates.size(); qa.num_selectors = k->num_selectors; qa.num_constants = NC;
        qa.num_routed = R; qa.num_wires = W; qa.num_challenges = nch; qa.npp = npp; qa.qdf = qdf; qa.num_gate_constraints = k->num_gate_constraints;
        qa.log_n = nlog; qa.rate_bits = rb;
        for (u32 c = 0; c < nch; c++) { qa.betas[c] = betas[c]; qa.gammas[c] = gammas[c]; }
        {
            std::vector<u64> bk((size_t)nch * R);
            for (u32 c = 0; c < nch; c++) for (u32 j = 0; j < R; j++) bk[(size_t)c * R + j] = gl_mul(betas[c], k->k_is[j]);
            u64 *d_bk = A.take(bk.size() + 8);
            S.h2d(d_bk, bk.data(), bk.size() * 8);
            qa.beta_kis = d_bk;
        }
        for (int i = 0; i < 4; i++) qa.pi_hash[i] = pi_hash[i];
        {
            u64 g_pow_n = gl_pow(GL_GEN, (u64)n), wr = gl_root_of_unity(rb), x = 1;
            for (u32 i = 0; i < (1u << rb); i++) { qa.zh_ev[i] = gl_sub(gl_mul(g_pow_n, x), 1); qa.zh_inv[i] = gl_inv(qa.zh_ev[i]); x = gl_mul(x, wr); }
            qa.n_inv = gl_inv((u64)n % GL_P);
        }
        qa.acc = q_big;     // scratch until the iNTT overwrites it
        // low-degree gates are evaluated on a sub-coset and extended (plonk.cu: k_gate / k_gate_combine)
        const u64 *unit_scale[4] = {nullptr, nullptr, nullptr, nullptr};
        u64 *d_sub = nullptr;
        if (sub_scratch_u64) {
            d_sub = A.take(sub_scratch_u64);
            for (u32 gi = 0; gi < k->gates.size(); gi++) {
                const u32 s = quotient_gate_sub_shift(k->gates[gi], nlog, rb);
                if (s && !unit_scale[s]) unit_scale[s] = ctx_lde_scale_unit(ctx, Llog - (int)s, (int)s);
            }
        }
        { PhaseTimer t(ctx, "quotient"); launch_quotient(qa, k->gates.data(), ctx->tw, unit_scale, d_sub, st); }
        // coefficients of Q(gX): DIF inverse NTT natural -> bit-reversed, scaled by 1/N
        { PhaseTimer t(ctx, "quotient_intt"); ntt_dif(ctx->tw, q_vals, N, false, q_big, N, Llog, nch, 0, nullptr, true, gl_inv((u64)N % GL_P), st); }
        // q_vals is free again: reuse as chunk coefficient storage?  keep separate buffers for clarity
        (void)q_vals;
        ctx->tmp_q_big = q_big;
    }
    u64 *q_coeffs = A.take((size_t)nq * n), *q_lde = A.take((size_t)nq * N), *q_dig = A.take(dig_u64);
    launch_quotient_split(ctx->tmp_q_big, q_coeffs, nlog, rb, nch, A.take(quotient_split_scratch_u64(nlog)), st);
    commit_coeffs(ctx, q_coeffs, nq, nlog, rb, chh, q_lde, q_dig, HS);
    S.d2h(cap_q.data(), q_dig + dig_u64 - capw, capw * 8);
    S.sync();
    ch.observe_cap(cap_q.data(), ncap, HS);
    gl2 zeta = ch.get_ext();
    if (gl2_eq(gl2_pow(zeta, (u64)n), gl2_make(1, 0))) throw ApiErr{B200_ERR_UNSATISFIED, "opening point zeta lies in the subgroup"};

    // ---- openings ----
    const u64 w_n = gl_root_of_unity(nlog);
    gl2 gzeta = gl2_scale(zeta, w_n);
    const size_t total_polys = (size_t)NC + R + W + nzp + nq;
    gl2 *d_tab = (gl2 *)A.take(2 * n), *d_tab2 = (gl2 *)A.take(2 * n);
    gl2 *d_open = (gl2 *)A.take(2 * (total_polys + nch));
    gl2 *d_open_part = (gl2 *)A.take(2 * eval_polys_scratch_gl2(std::max<size_t>(std::max<size_t>(W, NC + R), std::max<size_t>(nzp, nq))));
    {
        PhaseTimer tm(ctx, "openings");
        gl2 p2[32]; gl2 t = zeta; for (int b = 0; b < nlog; b++) { p2[b] = t; t = gl2_mul(t, t); }
        launch_ext_pow_table(d_tab, nlog, p2, st);
        t = gzeta; for (int b = 0; b < nlog; b++) { p2[b] = t; t = gl2_mul(t, t); }
        launch_ext_pow_table(d_tab2, nlog, p2, st);
        size_t off = 0;
        launch_eval_polys(k->cs_coeffs.u(), NC + R, nlog, d_tab, d_open + off, d_open_part, st); off += NC + R;
        launch_eval_polys(w_coeffs, W, nlog, d_tab, d_open + off, d_open_part, st); off += W;
        launch_eval_polys(zs_coeffs, nzp, nlog, d_tab, d_open + off, d_open_part, st); off += nzp;
        launch_eval_polys(q_coeffs, nq, nlog, d_tab, d_open + off, d_open_part, st); off += nq;
        launch_eval_polys(zs_coeffs, nch, nlog, d_tab2, d_open + off, d_open_part, st);
    }
    std::vector<gl2> open(total_polys + nch);
    S.d2h(open.data(), d_open, open.size() * sizeof(gl2));
    S.sync();
    for (size_t i = 0; i < total_polys; i++) ch.observe_ext(open[i]);
    for (u32 i = 0; i < nch; i++) ch.observe_ext(open[total_polys + i]);

    // ---- FRI ----
    gl2 fri_alpha = ch.get_ext();
    std::vector<gl2> apow(total_polys);
    { gl2 p = gl2_make(1, 0); for (size_t i = 0; i < total_polys; i++) { apow[i] = p; p = gl2_mul(p, fri_alpha); } }
    gl2 y0 = gl2_make(0, 0), y1 = gl2_make(0, 0);
    for (size_t i = 0; i < total_polys; i++) y0 = gl2_add(y0, gl2_mul(open[i], apow[i]));
    for (u32 i = 0; i < nch; i++) y1 = gl2_add(y1, gl2_mul(open[total_polys + i], apow[i]));
    gl2 *d_apow2 = (gl2 *)A.take(2 * total_polys);
    S.h2d(d_apow2, apow.data(), total_polys * sizeof(gl2));
    gl2 *fri_a = (gl2 *)A.take(2 * N), *fri_b = (gl2 *)A.take(2 * N);
    {
        FriCombineArgs fa; memset(&fa, 0, sizeof fa);
        fa.coeffs[0] = k->cs_coeffs.u(); fa.coeffs[1] = w_coeffs; fa.coeffs[2] = zs_coeffs; fa.coeffs[3] = q_coeffs;
        fa.cols[0] = NC + R; fa.cols[1] = W; fa.cols[2] = nzp; fa.cols[3] = nq;
        fa.apow = d_apow2; fa.total_polys = (u32)total_polys; fa.nch = nch; fa.y0 = y0; fa.y1 = y1; fa.zeta = zeta; fa.gzeta = gzeta;
        fa.alpha_nch = gl2_pow(fri_alpha, nch); fa.tw_fwd = ctx->tw.fwd; fa.log_n = nlog; fa.log_N = Llog; fa.out = fri_a;
        u64 *g = A.take(4 * n), *g_lde = A.take(4 * N);
        fa.g = g; fa.g_lde = g_lde;
        PhaseTimer t(ctx, "fri_combine");
        launch_fri_coeff_combine(fa, st);
        ntt_dif(ctx->tw, g, n, true, g_lde, N, nlog, (size_t)4 << rb, rb, ctx_lde_scale(ctx, nlog, rb), false, 1, st);
        launch_fri_combine(fa, st);
    }
    // The prover checks its own openings: a violated GATE constraint leaves every committed polynomial low-degree (the
    // quotient is interpolated from its coset values), so FRI cannot see it -- only the identity at zeta can.  The check is
    // host work (every gate evaluated once over the extension field); it runs while the device is busy with the FRI
    // combine and the first layer's tree.
    bool self_checked = false;
    auto self_check = [&]() {
        if (self_checked) return;
        self_checked = true;
        const gl2 *o = open.data();
        VanishingInput vi{W, R, NC, nch, qdf, npp, k->num_gate_constraints, k->num_selectors, nlog, k->k_is.data(), &k->gates, k->extra.data(), pi_hash,
                          betas, gammas, alphas, zeta, o, o + NC, o + NC + R, o + NC + R + W, o + total_polys, o + NC + R + W + nch, o + NC + R + W + nzp};
        if (const char *why = vanishing_identity_at_zeta(vi))
            throw ApiErr{B200_ERR_UNSATISFIED, std::string("witness does not satisfy the circuit: ") + why + " (a gate constraint is violated)"};
    };
    // commit phase: layers[l] = values (device, kept for the query phase), digests
    struct Layer { gl2 *vals; u64 *dig; int log_len, log_leaves; size_t dig_u64; };
    std::vector<Layer> layers;
    std::vector<std::vector<u64>> layer_caps;
    std::vector<gl2> fri_betas;
    gl2 *cur = fri_a; int cur_log = Llog; u64 shift = GL_GEN;
    size_t fri_pool_used = 0; gl2 *pool = fri_b;
    for (size_t li = 0; li < k->arity_bits.size(); li++) {
        int ab = (int)k->arity_bits[li];
        PhaseTimer tl(ctx, "fri_layers");
        Layer Ly; Ly.vals = cur; Ly.log_len = cur_log; Ly.log_leaves = cur_log - ab; Ly.dig_u64 = merkle_num_u64(Ly.log_leaves, chh);
        Ly.dig = A.take(Ly.dig_u64);
        if (HS == HASHER_BN254) { merkle_leaves_rowmajor_bn254((const u64 *)cur, size_t(2) << ab, Ly.log_leaves, Ly.dig, st); merkle_build_bn254(Ly.dig, Ly.log_leaves, chh, st); }
        else { merkle_leaves_rowmajor((const u64 *)cur, size_t(2) << ab, Ly.log_leaves, Ly.dig, st); merkle_build(Ly.dig, Ly.log_leaves, chh, st); }
        std::vector<u64> cap(capw);
        S.d2h(cap.data(), Ly.dig + Ly.dig_u64 - capw, capw * 8);
        self_check();
        S.sync();
        ch.observe_cap(cap.data(), ncap, HS);
        gl2 beta = ch.get_ext();
        fri_betas.push_back(beta); layer_caps.push_back(cap); layers.push_back(Ly);
        gl2 *nxt = pool + fri_pool_used; fri_pool_used += size_t(1) << (cur_log - ab);
        launch_fri_fold(cur, nxt, cur_log, ab, beta, shift, ctx->tw.fwd, st);
        shift = gl_pow(shift, u64(1) << ab);
        cur = nxt; cur_log -= ab;
    }
    // final polynomial (host): values on shift * H_{2^cur_log} in bit-reversed order -> coefficients
    size_t fin_n = size_t(1) << cur_log;
    std::vector<gl2> fin_vals(fin_n);
    S.d2h(fin_vals.data(), cur, fin_n * sizeof(gl2));
    self_check();
    S.sync();
    std::vector<u64> fa(fin_n), fb(fin_n);
    for (size_t i = 0; i < fin_n; i++) { size_t r = bitrev32((u32)i, cur_log); fa[i] = fin_vals[r].a; fb[i] = fin_vals[r].b; }
    host_ntt(fa, cur_log, true); host_ntt(fb, cur_log, true);
    { u64 sinv = gl_inv(shift), s = 1; for (size_t i = 0; i < fin_n; i++) { fa[i] = gl_mul(fa[i], s); fb[i] = gl_mul(fb[i], s); s = gl_mul(s, sinv); } }
    size_t fin_len = fin_n >> rb;
    for (size_t i = fin_len; i < fin_n; i++) if (fa[i] || fb[i]) throw ApiErr{B200_ERR_UNSATISFIED, "witness does not satisfy the circuit: FRI final polynomial is not low-degree"};
    for (size_t i = 0; i < fin_len; i++) { ch.observe(fa[i]); ch.observe(fb[i]); }

    // ---- proof of work: minimum nonce (== reference with RAYON_NUM_THREADS=1), or a verified hint ----
    u64 pow_witness = ~u64(0);
    {
        u64 base_state[12]; memcpy(base_state, ch.st, sizeof base_state);
        for (int i = 0; i < ch.nin; i++) base_state[i] = ch.in[i];
        int pos = ch.nin;
        auto response = [&](u64 cand) { u64 s[12]; memcpy(s, base_state, sizeof s); s[pos] = cand; poseidon_permute(s); return s[7]; };
        if (pow_hint && pow_hint < GL_P && (response(pow_hint) >> (64 - k->pow_bits)) == 0) pow_witness = pow_hint;
        PhaseTimer tp(ctx, "pow");
        u64 *d_res = A.take(8);
        u64 start = 0; u64 chunk = u64(1) << (k->pow_bits + 1 < 24 ? k->pow_bits + 1 : 24);   // ~86 % hit rate per first launch
        while (pow_witness == ~u64(0)) {
            u64 init = ~u64(0);
            S.h2d(d_res, &init, 8);
            launch_pow(base_state, pos, start, chunk, (int)k->pow_bits, d_res, st);
            u64 res; S.d2h(&res, d_res, 8);
            S.sync();
            if (res != ~u64(0)) pow_witness = res;
            start += chunk;
            if (chunk < (u64(1) << 22)) chunk <<= 1;
            if (start > (u64(1) << 40)) throw ApiErr{B200_ERR_CUDA, "proof of work not found"};
        }
        ch.observe(pow_witness);
        u64 resp = ch.get();
        if (resp >> (64 - k->pow_bits)) throw ApiErr{B200_ERR_CUDA, "proof-of-work self-check failed"};
    }

    // ---- query rounds: gather leaves + Merkle paths with one kernel ----
    const u64 *lde4[4] = {k->cs_lde.u(), w_lde, zs_lde, q_lde};
    const u64 *dig4[4] = {k->cs_digests.u(), w_dig, zs_dig, q_dig};
    const u32 cols4[4] = {NC + R, W, nzp, nq};
    std::vector<CopyDesc> descs;
    std::vector<size_t> qidx(k->num_queries);
    u32 dst = 0;
    for (u32 q = 0; q < k->num_queries; q++) {
        size_t x = (size_t)(ch.get() % (u64)N);
        qidx[q] = x;
        for (int o = 0; o < 4; o++) {
            descs.push_back({lde4[o] + x, (u64)N, cols4[o], dst}); dst += cols4[o];
            const u64 *lvl = dig4[o]; size_t idx = x;
            for (int lv = Llog; lv > chh; lv--) { descs.push_back({lvl + 4 * (idx ^ 1), 1, 4, dst}); dst += 4; lvl += size_t(4) << lv; idx >>= 1; }
        }
        size_t xi = x;
        for (size_t li = 0; li < layers.size(); li++) {
            int ab = (int)k->arity_bits[li]; size_t leaf = xi >> ab;
            descs.push_back({(const u64 *)layers[li].vals + leaf * (size_t(2) << ab), 1, (u32)(2u << ab), dst}); dst += 2u << ab;
            const u64 *lvl = layers[li].dig; size_t idx = leaf;
            for (int lv = layers[li].log_leaves; lv > chh; lv--) { descs.push_back({lvl + 4 * (idx ^ 1), 1, 4, dst}); dst += 4; lvl += size_t(4) << lv; idx >>= 1; }
            xi = leaf;
        }
    }
    std::vector<u64> qdata(dst);
    if (!descs.empty()) {
        PhaseTimer tq(ctx, "query_gather");
        CopyDesc *d_descs = (CopyDesc *)A.take(descs.size() * sizeof(CopyDesc) / 8 + 8);
        u64 *d_q = A.take(dst + 8);
        S.h2d(d_descs, descs.data(), descs.size() * sizeof(CopyDesc));
        k_copy_descs<<<(unsigned)descs.size(), 64, 0, st>>>(d_descs, d_q);
        LAUNCHED(1);
        S.d2h(qdata.data(), d_q, (size_t)dst * 8);
        S.sync();
    }

    // ---- serialize (ProofWithPublicInputs::to_bytes layout) ----
    std::vector<uint8_t> out;
    out.reserve(200000);
    put_words(out, cap_w.data(), capw); put_words(out, cap_z.data(), capw); put_words(out, cap_q.data(), capw);
    {
        size_t o_sig = NC, o_w = NC + R, o_z = o_w + W, o_pp = o_z + nch, o_q = o_z + nzp;
        static_assert(sizeof(gl2) == 16, "gl2 is two packed u64 words");
        auto put_ext = [&](size_t a, size_t b) { put_words(out, (const u64 *)(open.data() + a), 2 * (b - a)); };
        put_ext(0, o_sig); put_ext(o_sig, o_w); put_ext(o_w, o_z); put_ext(o_z, o_pp);
        put_ext(total_polys, total_polys + nch);          // plonk_zs_next
        put_ext(o_pp, o_q); put_ext(o_q, total_polys);
    }
    for (auto &c : layer_caps) put_words(out, c.data(), capw);
    {
        size_t p = 0;
        for (u32 q = 0; q < k->num_queries; q++) {
            for (int o = 0; o < 4; o++) {
                put_words(out, qdata.data() + p, cols4[o]); p += cols4[o];
                int ns = Llog - chh; out.push_back((uint8_t)ns);
                put_words(out, qdata.data() + p, (size_t)4 * ns); p += (size_t)4 * ns;
            }
            for (size_t li = 0; li < layers.size(); li++) {
                size_t w = size_t(2) << k->arity_bits[li];
                put_words(out, qdata.data() + p, w); p += w;
                int ns = layers[li].log_leaves - chh; out.push_back((uint8_t)ns);
                put_words(out, qdata.data() + p, (size_t)4 * ns); p += (size_t)4 * ns;
            }
        }
    }
    for (size_t i = 0; i < fin_len; i++) { put_u64(out, fa[i]); put_u64(out, fb[i]); }
    put_u64(out, pow_witness);
    put_u64(out, (u64)n_pis);
    put_words(out, pis, n_pis);

    uint8_t *buf = (uint8_t *)malloc(out.size());
    if (!buf) throw ApiErr{B200_ERR_OOM, "host allocation failed"};
    memcpy(buf, out.data(), out.size());
    *proof_out = buf; *proof_len = out.size();
    return B200_OK;
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

// ---- shared by b200_circuit_upload and b200_circuit_build: validate `c`, parse the gate table ----
static std::unique_ptr<b200_circuit> circuit_from_common(b200_ctx *ctx, const b200_common *c) {
    if (c->hasher > 1) throw ApiErr{B200_ERR_UNSUPPORTED, "hasher must be 0 (PoseidonGoldilocksConfig) or 1 (PoseidonBN128GoldilocksConfig)"};
    if (c->num_challenges == 0 || c->num_challenges > MAX_CHALLENGES) throw ApiErr{B200_ERR_UNSUPPORTED, "num_challenges must be 1 or 2"};
    if (c->rate_bits == 0 || c->rate_bits > 4) throw ApiErr{B200_ERR_UNSUPPORTED, "rate_bits must be in 1..4"};     // QuotientArgs::zh_ev[16]
    if (c->quotient_degree_factor != (1u << c->rate_bits)) throw ApiErr{B200_ERR_UNSUPPORTED, "quotient_degree_factor must equal 2^rate_bits"};
    if (c->num_reduction_arity_bits > 16) throw ApiErr{B200_ERR_BAD_ARG, "too many FRI reductions"};
    if (c->degree_bits < 2 || c->degree_bits + c->rate_bits > 28) throw ApiErr{B200_ERR_BAD_ARG, "degree_bits out of range"};
    if (c->num_routed_wires > c->num_wires || c->num_partial_products + 1 > 16 ||
        (c->num_partial_products + 1) * c->quotient_degree_factor < c->num_routed_wires)
        throw ApiErr{B200_ERR_BAD_ARG, "inconsistent routed wires / partial products"};
    if (c->cap_height > c->degree_bits + c->rate_bits || c->proof_of_work_bits == 0 || c->proof_of_work_bits > 32)
        throw ApiErr{B200_ERR_BAD_ARG, "bad FRI config"};
    std::unique_ptr<b200_circuit> k(new b200_circuit());
    k->ctx = ctx;
    k->num_wires = c->num_wires; k->num_routed = c->num_routed_wires; k->num_constants = c->num_constants; k->num_challenges = c->num_challenges;
    k->degree_bits = c->degree_bits; k->rate_bits = c->rate_bits; k->cap_height = c->cap_height; k->pow_bits = c->proof_of_work_bits;
    k->num_queries = c->num_query_rounds; k->qdf = c->quotient_degree_factor; k->npp = c->num_partial_products;
    k->num_gate_constraints = c->num_gate_constraints; k->num_pi = c->num_public_inputs; k->num_selectors = c->num_selector_groups; k->hasher = c->hasher;
    int bits_left = (int)c->degree_bits + (int)c->rate_bits;
    for (u32 i = 0; i < c->num_reduction_arity_bits; i++) {
        u32 ab = c->reduction_arity_bits[i];
        if (ab < 1 || ab > 4 || bits_left - (int)ab < (int)c->cap_height) throw ApiErr{B200_ERR_UNSUPPORTED, "reduction arity bits must be in 1..4 and leave >= cap_height bits"};
        k->arity_bits.push_back(ab); bits_left -= (int)ab;
    }
    if (bits_left < (int)c->rate_bits) throw ApiErr{B200_ERR_BAD_ARG, "too many FRI reductions"};
    std::vector<u64> &extra = k->extra;
    for (u32 i = 0; i < c->num_gates; i++) {
        GateDesc g;
        if (!c->gate_ids[i] || !parse_gate(c->gate_ids[i], g, extra)) { std::string m = "unsupported gate id: "; m += c->gate_ids[i] ? c->gate_ids[i] : "(null)"; throw ApiErr{B200_ERR_UNSUPPORTED, m}; }
        u32 si = c->selector_indices[i];
        if (si >= c->num_selector_groups) throw ApiErr{B200_ERR_BAD_ARG, "selector index out of range"};
        g.selector_index = si; g.group_start = c->group_starts[si]; g.group_end = c->group_ends[si]; g.row = i;
        if (g.num_constraints > c->num_gate_constraints) throw ApiErr{B200_ERR_BAD_ARG, "num_gate_constraints too small"};
        k->gates.push_back(g);
    }
    k->k_is.assign(c->k_is, c->k_is + c->num_routed_wires);
    return k;
}
// commit device-resident constants/sigmas VALUES [num_constants + num_routed][n] (consumed), derive the digest if none is given
static void circuit_commit(b200_ctx *ctx, b200_circuit *k, DevBuf &vals, const uint64_t *digest, uint64_t *cap_out) {
    cudaStream_t st = ctx->stream;
    const size_t n = size_t(1) << k->degree_bits, N = n << k->rate_bits;
    const size_t cols = (size_t)k->num_constants + k->num_routed;
    k->d_gates.alloc(std::max<size_t>(1, k->gates.size()) * sizeof(GateDesc));
    k->d_extra.alloc(std::max<size_t>(1, k->extra.size()) * 8);
    k->d_kis.alloc(std::max<size_t>(1, k->k_is.size()) * 8);
    CUDA_TRY(cudaMemcpyAsync(k->d_gates.p, k->gates.data(), k->gates.size() * sizeof(GateDesc), cudaMemcpyHostToDevice, st));
    if (!k->extra.empty()) CUDA_TRY(cudaMemcpyAsync(k->d_extra.p, k->extra.data(), k->extra.size() * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(k->d_kis.p, k->k_is.data(), k->k_is.size() * 8, cudaMemcpyHostToDevice, st));
    k->cs_coeffs.alloc(cols * n * 8); k->cs_lde.alloc(cols * N * 8);
    const size_t nd = merkle_num_u64((int)(k->degree_bits + k->rate_bits), (int)k->cap_height);
    k->cs_digests.alloc(nd * 8);
    commit_values(ctx, vals.u(), cols, (int)k->degree_bits, (int)k->rate_bits, (int)k->cap_height, k->cs_coeffs.u(), k->cs_lde.u(), k->cs_digests.u(), (int)k->hasher);
    k->sigma_values.alloc((size_t)k->num_routed * n * 8);
    CUDA_TRY(cudaMemcpyAsync(k->sigma_values.p, vals.u() + (size_t)k->num_constants * n, (size_t)k->num_routed * n * 8, cudaMemcpyDeviceToDevice, st));
    const size_t capw = size_t(4) << k->cap_height;
    k->cap.resize(capw);
    CUDA_TRY(cudaMemcpyAsync(k->cap.data(), k->cs_digests.u() + nd - capw, capw * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (digest) memcpy(k->digest, digest, 32);
    else circuit_digest_host(k->cap.data(), size_t(1) << k->cap_height, k->degree_bits, (int)k->hasher, k->digest);
    if (cap_out) memcpy(cap_out, k->cap.data(), capw * 8);
}

// ---- CircuitBuilder::build()'s device half -------------------------------------------------------------------------
// selector columns from the gate placement: column g holds the gate index on rows whose gate belongs to selector group g,
// UNUSED (2^32 - 1) elsewhere (a single group: just the index); reference semantics: SURVEY.md A.2 / evaluate_gates.go:33-55
__global__ void k_build_selectors(const u32 *__restrict__ row_gate, const u32 *__restrict__ gate_group, u32 nsel, size_t n, u64 *__restrict__ cols) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 g = row_gate[i], grp = gate_group[g];
    for (u32 s = 0; s < nsel; s++) cols[(size_t)s * n + i] = (s == grp) ? (u64)g : (nsel > 1 ? 0xFFFFFFFFULL : 0ULL);
}
// sigma polynomial values: slot (col, row) -> k_is[col'] * w^row' of the NEXT wire (col', row') of its copy class
__global__ void k_build_sigmas(const u32 *__restrict__ sigma_next, const u64 *__restrict__ kis, const u64 *__restrict__ tw, int log_n, u32 R, u64 *__restrict__ sig) {
    const size_t n = size_t(1) << log_n, total = (size_t)R << log_n;
    size_t s = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (s >= total) return;
    const u32 nx = sigma_next[s];
    const u32 col = nx >> log_n; const size_t row = nx & (n - 1);
    const size_t h = n >> 1;
    u64 x = (row < h) ? tw[h + row] : gl_neg(tw[h + (row - h)]);      // w_n^row from the twiddle table tw[h + j] = w_{2h}^j
    sig[s] = gl_mul(kis[col], x);
}
// 128-bit content key of the build inputs (not cryptographic: the inputs are the caller's own circuit, not adversarial)
static void key_mix(u64 key[2], const void *data, size_t bytes) {
    const u64 *w = (const u64 *)data; size_t nw = bytes / 8;
    u64 a = key[0], b = key[1];
    for (size_t i = 0; i < nw; i++) { u64 x = w[i]; a = (a ^ x) * 0x9E3779B97F4A7C15ULL; a ^= a >> 29; b = (b + x) * 0xC2B2AE3D27D4EB4FULL; b ^= b >> 31; }
    const uint8_t *t = (const uint8_t *)data + nw * 8;
    for (size_t i = 0; i < bytes % 8; i++) { a = (a ^ t[i]) * 0x100000001B3ULL; b = (b + t[i]) * 0x9E3779B97F4A7C15ULL; }
    key[0] = a; key[1] = b ^ (a >> 32);
}

void ctx_drop_circuit_cache(b200_ctx *ctx) {
    for (b200_circuit *cc : ctx->circuit_cache) delete cc;
    ctx->circuit_cache.clear();
}

extern "C" {

int b200_circuit_upload(b200_ctx *ctx, const b200_common *c, const uint64_t *cs_values, const uint64_t *digest, uint64_t *cap_out, b200_circuit **out) {
    API_BEGIN(ctx)
    if (!c || !cs_values || !out) throw ApiErr{B200_ERR_BAD_ARG, "b200_circuit_upload: null argument"};
    *out = nullptr;
    std::unique_ptr<b200_circuit> k = circuit_from_common(ctx, c);
    const size_t n = size_t(1) << c->degree_bits;
    const size_t cols = (size_t)c->num_constants + c->num_routed_wires;
    DevBuf vals(cols * n * 8);
    CUDA_TRY(cudaMemcpyAsync(vals.p, cs_values, cols * n * 8, cudaMemcpyHostToDevice, ctx->stream));
    circuit_commit(ctx, k.get(), vals, digest, cap_out);
    *out = k.release();
    return B200_OK;
    API_END(ctx)
}

int b200_circuit_build(b200_ctx *ctx, const b200_common *c, const uint32_t *row_gate, const uint64_t *gate_constants, const uint32_t *sigma_next,
                       uint64_t *cap_out, uint64_t *digest_out, int *cache_hit, b200_circuit **out) {
    API_BEGIN(ctx)
    if (!c || !row_gate || !sigma_next || !out) throw ApiErr{B200_ERR_BAD_ARG, "b200_circuit_build: null argument"};
    *out = nullptr;
    if (c->num_constants < c->num_selector_groups) throw ApiErr{B200_ERR_BAD_ARG, "num_constants smaller than the number of selector groups"};
    const u32 ngc = c->num_constants - c->num_selector_groups;
    if (ngc && !gate_constants) throw ApiErr{B200_ERR_BAD_ARG, "b200_circuit_build: gate constants missing"};
    if (c->degree_bits > 26) throw ApiErr{B200_ERR_BAD_ARG, "degree_bits out of range"};
    const size_t n = size_t(1) << c->degree_bits;
    // ---- cache lookup: same gate table, placement, constants and routing => same CircuitData (recursion.rs:36-96 rebuilds
    // an identical recursion circuit on every one of its k + 55 calls) ----
    u64 key[2] = {0x243F6A8885A308D3ULL, 0x13198A2E03707344ULL};
    {
        u32 hdr[16] = {c->num_wires, c->num_routed_wires, c->num_constants, c->num_challenges, c->degree_bits, c->rate_bits, c->cap_height,
                       c->proof_of_work_bits, c->num_query_rounds, c->quotient_degree_factor, c->num_partial_products, c->num_gate_constraints,
                       c->num_public_inputs, c->hasher, c->num_gates, c->num_selector_groups};
        key_mix(key, hdr, sizeof hdr);
        key_mix(key, c->reduction_arity_bits, sizeof(u32) * std::min<u32>(c->num_reduction_arity_bits, 16));
        for (u32 i = 0; i < c->num_gates; i++) if (c->gate_ids[i]) key_mix(key, c->gate_ids[i], strlen(c->gate_ids[i]));
        key_mix(key, c->selector_indices, sizeof(u32) * c->num_gates);
        key_mix(key, c->group_starts, sizeof(u32) * c->num_selector_groups); key_mix(key, c->group_ends, sizeof(u32) * c->num_selector_groups);
        key_mix(key, c->k_is, 8 * (size_t)c->num_routed_wires);
        key_mix(key, row_gate, 4 * n);
        if (ngc) key_mix(key, gate_constants, 8 * n * ngc);
        key_mix(key, sigma_next, 4 * n * c->num_routed_wires);
    }
    ctx->cache_tick++;
    for (b200_circuit *cc : ctx->circuit_cache) if (cc->key[0] == key[0] && cc->key[1] == key[1]) {
        cc->refs++; cc->last_use = ctx->cache_tick; ctx->cache_hits++;
        if (cap_out) memcpy(cap_out, cc->cap.data(), cc->cap.size() * 8);
        if (digest_out) memcpy(digest_out, cc->digest, 32);
        if (cache_hit) *cache_hit = 1;
        *out = cc;
        return B200_OK;
    }
    ctx->cache_misses++;
    if (cache_hit) *cache_hit = 0;
    std::unique_ptr<b200_circuit> k = circuit_from_common(ctx, c);
    for (u32 i = 0; i < n; i++) if (row_gate[i] >= c->num_gates) throw ApiErr{B200_ERR_BAD_ARG, "row_gate entry out of range"};
    cudaStream_t st = ctx->stream;
    const u32 R = c->num_routed_wires, nsel = c->num_selector_groups;
    if (((u64)R << c->degree_bits) > 0xFFFFFFFFULL) throw ApiErr{B200_ERR_UNSUPPORTED, "routed slots do not fit 32-bit indices"};
    ntt_tables_ensure(ctx->tw, (int)c->degree_bits + (int)c->rate_bits, st);
    const size_t cols = (size_t)c->num_constants + R;
    DevBuf vals(cols * n * 8), d_row(4 * n), d_grp(4 * std::max<u32>(1, c->num_gates)), d_next(4 * n * R), d_kis(8 * std::max<u32>(1, R));
    CUDA_TRY(cudaMemcpyAsync(d_row.p, row_gate, 4 * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_grp.p, c->selector_indices, 4 * (size_t)c->num_gates, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_next.p, sigma_next, 4 * n * R, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_kis.p, c->k_is, 8 * (size_t)R, cudaMemcpyHostToDevice, st));
    if (ngc) CUDA_TRY(cudaMemcpyAsync(vals.u() + (size_t)nsel * n, gate_constants, 8 * n * ngc, cudaMemcpyHostToDevice, st));
    {
        PhaseTimer t(ctx, "build_polys");
        k_build_selectors<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const u32 *)d_row.p, (const u32 *)d_grp.p, nsel, n, vals.u());
        LAUNCHED(1);
        k_build_sigmas<<<(unsigned)((n * R + 255) / 256), 256, 0, st>>>((const u32 *)d_next.p, d_kis.u(), ctx->tw.fwd, (int)c->degree_bits, R, vals.u() + (size_t)c->num_constants * n);
        LAUNCHED(1);
    }
    circuit_commit(ctx, k.get(), vals, nullptr, cap_out);
    if (digest_out) memcpy(digest_out, k->digest, 32);
    // ---- insert into the cache (the cache holds one reference); evict unreferenced entries beyond the capacity, oldest first
    k->cached = true; k->key[0] = key[0]; k->key[1] = key[1]; k->last_use = ctx->cache_tick; k->refs = 2;
    b200_circuit *raw = k.release();
    ctx->circuit_cache.push_back(raw);
    while (ctx->circuit_cache.size() > 8) {
        size_t victim = ctx->circuit_cache.size();
        for (size_t i = 0; i < ctx->circuit_cache.size(); i++)
            if (ctx->circuit_cache[i]->refs == 1 && (victim == ctx->circuit_cache.size() || ctx->circuit_cache[i]->last_use < ctx->circuit_cache[victim]->last_use)) victim = i;
        if (victim == ctx->circuit_cache.size()) break;             // everything is in use
        delete ctx->circuit_cache[victim];
        ctx->circuit_cache.erase(ctx->circuit_cache.begin() + victim);
    }
    *out = raw;
    return B200_OK;
    API_END(ctx)
}

int b200_circuit_digest(uint32_t hasher, const uint64_t *cap, uint32_t cap_height, uint32_t degree_bits, uint64_t out[4]) {
    if (hasher > 1 || !cap || !out || cap_height > 16) return B200_ERR_BAD_ARG;
    circuit_digest_host(cap, size_t(1) << cap_height, degree_bits, (int)hasher, out);
    return B200_OK;
}

int b200_ctx_cache_stats(const b200_ctx *ctx, uint64_t *hits, uint64_t *misses, uint64_t *entries) {
    if (!ctx) return B200_ERR_BAD_ARG;
    if (hits) *hits = ctx->cache_hits; if (misses) *misses = ctx->cache_misses; if (entries) *entries = ctx->circuit_cache.size();
    return B200_OK;
}


void b200_circuit_free(b200_circuit *c) {
    if (!c) return;
    if (c->cached) { if (c->refs > 1) c->refs--; return; }       // the context's cache keeps it (freed with the context or on eviction)
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    delete c;
}

int b200_prove(b200_ctx *ctx, const b200_circuit *circuit, const uint64_t *wires, const uint64_t *public_inputs, size_t n_pis,
               uint64_t pow_witness_hint, uint8_t **proof_bytes, size_t *proof_len) {
    API_BEGIN(ctx)
    if (!circuit || !wires || !proof_bytes || !proof_len || (n_pis && !public_inputs)) throw ApiErr{B200_ERR_BAD_ARG, "b200_prove: null argument"};
    if (circuit->ctx != ctx) throw ApiErr{B200_ERR_BAD_ARG, "b200_prove: circuit belongs to another context"};
    return prove_impl(ctx, circuit, wires, false, public_inputs, n_pis, pow_witness_hint, proof_bytes, proof_len);
    API_END(ctx)
}

int b200_prove_dev(b200_ctx *ctx, const b200_circuit *circuit, const uint64_t *d_wires, const uint64_t *public_inputs, size_t n_pis,
                   uint64_t pow_witness_hint, uint8_t **proof_bytes, size_t *proof_len) {
    API_BEGIN(ctx)
    if (!circuit || !d_wires || !proof_bytes || !proof_len || (n_pis && !public_inputs)) throw ApiErr{B200_ERR_BAD_ARG, "b200_prove_dev: null argument"};
    if (circuit->ctx != ctx) throw ApiErr{B200_ERR_BAD_ARG, "b200_prove_dev: circuit belongs to another context"};
    return prove_impl(ctx, circuit, d_wires, true, public_inputs, n_pis, pow_witness_hint, proof_bytes, proof_len);
    API_END(ctx)
}

int b200_fri_fold(b200_ctx *ctx, const uint64_t *values, uint32_t log_len, uint32_t arity_bits, const uint64_t beta[2], uint64_t shift, uint64_t *out) {
    API_BEGIN(ctx)
    if (!values || !beta || !out || arity_bits < 1 || arity_bits > 4 || log_len < arity_bits || log_len > 28) throw ApiErr{B200_ERR_BAD_ARG, "b200_fri_fold: bad arguments"};
    ntt_tables_ensure(ctx->tw, (int)log_len, ctx->stream);
    size_t len = size_t(1) << log_len, nout = len >> arity_bits;
    DevBuf a(len * 16), b(nout * 16);
    CUDA_TRY(cudaMemcpyAsync(a.p, values, len * 16, cudaMemcpyHostToDevice, ctx->stream));
    launch_fri_fold((const gl2 *)a.p, (gl2 *)b.p, (int)log_len, (int)arity_bits, gl2_make(beta[0], beta[1]), shift, ctx->tw.fwd, ctx->stream);
    CUDA_TRY(cudaMemcpyAsync(out, b.p, nout * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
    API_END(ctx)
}

}  // extern "C"
