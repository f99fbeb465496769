// Device-side building blocks of the prover beyond NTT/Merkle (declared here, defined in plonk.cu).
#pragma once
#include "b200_internal.h"
#include "gates.cuh"

struct QuotientArgs {
    const u64 *cs_lde, *wires_lde, *zs_lde;      // [cols][N] column-major, bit-reversed rows
    u64 *out;                                    // [num_challenges][N], written at the NATURAL point index
    u64 *acc;                                    // [num_challenges][N] scratch accumulator in storage (bit-reversed) order
    const GateDesc *gates; const u64 *extra; const u64 *k_is;
    const u64 *beta_kis;                         // [num_challenges][num_routed]: beta_c * k_j
    const u64 *apow;                             // [MAX_CHALLENGES][apow_stride] alpha powers
    const u64 *tw_fwd;                           // twiddle table (for x = g * w_N^i)
    u32 apow_stride;
    u32 num_gates, num_selectors, num_constants, num_routed, num_wires, num_challenges, npp, qdf, num_gate_constraints;
    int log_n, rate_bits;
    u64 betas[MAX_CHALLENGES], gammas[MAX_CHALLENGES];
    u64 pi_hash[4];
    u64 zh_inv[16], zh_ev[16];                   // indexed by i mod 2^rate_bits
    u64 n_inv;                                   // 1/n
};

// scratch words launch_partial_products needs: chunk quotients nch*(npp+1)*n, row products nch*n, scan block totals nch*ceil(n/1024)
static inline size_t partial_products_scratch_u64(int log_n, u32 npp, u32 nch) {
    const size_t n = size_t(1) << log_n;
    return (size_t)nch * (npp + 2) * n + (size_t)nch * ((n + 1023) / 1024) + 64;
}
void launch_partial_products(const u64 *d_wires, const u64 *d_sigmas, const u64 *d_kis, int log_n, u32 num_routed, u32 npp, u32 qdf,
                             u32 num_challenges, const u64 *betas, const u64 *gammas, const u64 *tw_fwd,
                             u64 *d_zs_pp /*[nch*(1+npp)][n]*/, u64 *d_scratch /* partial_products_scratch_u64() words */, u64 *d_final_prod /*[nch]*/,
                             cudaStream_t st);
// unit_scale[s] (s = 1..3): ctx_lde_scale_unit(log_n + rate_bits - s, s) or NULL; d_sub_scratch: quotient_sub_scratch_u64() words
// (NULL: every gate is evaluated on the whole coset)
void launch_quotient(const QuotientArgs &a, const GateDesc *h_gates, const NttTables &tw, const u64 *const unit_scale[4], u64 *d_sub_scratch, cudaStream_t st);
u32 quotient_gate_sub_shift(const GateDesc &g, int log_n, int rate_bits);
size_t quotient_sub_scratch_u64(const GateDesc *gates, u32 num_gates, int log_n, int rate_bits, u32 nch);
// big[ci][.] (size N, br order, coefficients of the coset-shifted quotient) -> chunk polys coeffs_br[ci*qdf + m][n]
void launch_quotient_split(const u64 *d_big, u64 *d_chunks, int log_n, int rate_bits, u32 num_challenges, u64 *d_scratch /* quotient_split_scratch_u64(log_n) words */, cudaStream_t st);
size_t quotient_split_scratch_u64(int log_n);
// table[q] = base^{bitrev(q)} (ext), q < 2^L ; pows2[b] = base^(2^b)
void launch_ext_pow_table(gl2 *d_table, int L, const gl2 *h_pows2, cudaStream_t st);
// out[p] = sum_q coeffs[p*n + q] * table[q]; d_partial: eval_polys_scratch_gl2(num_polys) elements of scratch
void launch_eval_polys(const u64 *d_coeffs_br, size_t num_polys, int L, const gl2 *d_table, gl2 *d_out, gl2 *d_partial, cudaStream_t st);
size_t eval_polys_scratch_gl2(size_t num_polys);
struct FriCombineArgs {
    const u64 *coeffs[4]; u32 cols[4];           // the four oracles' coefficient matrices [cols][n] (bit-reversed coefficient order)
    const gl2 *apow;                             // alpha^p for p < total_polys (device)
    u32 total_polys, nch;                        // zs oracle index is 2, its first nch columns are the Z polys
    gl2 y0, y1, zeta, gzeta, alpha_nch;          // reduced openings, points, alpha^{nch}
    const u64 *tw_fwd; int log_n, log_N;
    u64 *g;                                      // [4][n]: coefficients of G0.a, G0.b, G1.a, G1.b (launch_fri_coeff_combine writes)
    const u64 *g_lde;                            // [4][N]: their LDE on the commitment coset (launch_fri_combine reads)
    gl2 *out;                                    // [N] bit-reversed order
};
void launch_fri_coeff_combine(const FriCombineArgs &a, cudaStream_t st);
void launch_fri_combine(const FriCombineArgs &a, cudaStream_t st);
// fold: in [len] ext (br order), out [len >> arity_bits]
void launch_fri_fold(const gl2 *d_in, gl2 *d_out, int log_len, int arity_bits, gl2 beta, u64 shift, const u64 *tw_fwd, cudaStream_t st);
// proof of work: smallest candidate in [start, start+count) whose response has `bits` leading zeros, else ~0
void launch_pow(const u64 state[12], int pos, u64 start, u64 count, int bits, u64 *d_result, cudaStream_t st);
