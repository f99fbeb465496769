// Host emulation of the bulk-copy staged NTT passes (ntt2.cu): the same per-thread round functions, row <-> global mappings
// and twiddle builders (ntt2_core.cuh) run for tid = 0..255, round by round, over plain arrays; the asynchronous copies
// become memcpy of 16-element rows.  Checked against a textbook radix-2 NTT: forward / inverse transforms (natural in,
// bit-reversed storage out) and the x8 coset LDE from bit-reversed coefficient storage, for every supported size class.
// Built and run by tests/test_host_emulations.py (nvcc host compile, no device needed).
#include "ntt2_core.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
static u64 rnd() { return (((u64)rand() << 42) ^ ((u64)rand() << 21) ^ (u64)rand()) % GL_P; }
static std::vector<u64> make_tw(int max_log, bool inv) {
    std::vector<u64> t(size_t(1) << max_log, 0);
    for (int s = 0; s < max_log; s++) { u64 h = u64(1) << s, w = gl_root_of_unity(s + 1); if (inv) w = gl_inv(w); u64 x = 1; for (u64 j = 0; j < h; j++) { t[h + j] = x; x = gl_mul(x, w); } }
    return t;
}
static void ref_ntt(std::vector<u64> &a, int L, bool inv) {     // natural -> natural
    size_t n = size_t(1) << L;
    for (size_t i = 0; i < n; i++) { size_t j = bitrev32((u32)i, L); if (i < j) std::swap(a[i], a[j]); }
    u64 wn = gl_root_of_unity(L); if (inv) wn = gl_inv(wn);
    for (int s = 1; s <= L; s++) {
        size_t m = size_t(1) << s, h = m >> 1; u64 wm = gl_pow(wn, n >> s);
        std::vector<u64> tw(h); tw[0] = 1; for (size_t k = 1; k < h; k++) tw[k] = gl_mul(tw[k - 1], wm);
        for (size_t k = 0; k < n; k += m) for (size_t j = 0; j < h; j++) { u64 t = gl_mul(tw[j], a[k + j + h]), u = a[k + j]; a[k + j] = gl_add(u, t); a[k + j + h] = gl_sub(u, t); }
    }
}
template <int MODE>
static void run_pass(const N2Pass &p, const u64 *in, u64 *out, const u64 *tw, const u64 *scale, u64 out_mul) {
    const u32 tiles = 1u << (p.L - 12);
    std::vector<u64> IN(N2_BUF_WORDS), WORK(N2_BUF_WORDS), twB(256);
    for (u32 pos = 0; pos < tiles; pos++) {
        for (u32 r = 0; r < 256; r++) {
            u64 off = MODE == 1 ? n2_row_global_bitrev(p, pos, r) : n2_row_global(p, pos, r);
            memcpy(&IN[r * N2_ROW_WORDS], in + off, 128);
        }
        for (u32 t = 0; t < 240; t++) twB[(t / 15) * 16 + t % 15] = n2_twB_entry<MODE>(p, tw, pos, t / 15, t % 15);
        std::vector<u64> A(256 * 15, 0);
        for (u32 tid = 0; tid < 256; tid++) { u64 (&a)[15] = *reinterpret_cast<u64 (*)[15]>(&A[tid * 15]); n2_build_twsA<MODE>(p, tw, pos, tid, a); }
        for (u32 tid = 0; tid < 256; tid++) { u64 (&a)[15] = *reinterpret_cast<u64 (*)[15]>(&A[tid * 15]); n2_first_round<MODE>(p, IN.data(), MODE == 1 && scale ? scale + n2_row_global_bitrev(p, pos, tid) : nullptr, WORK.data(), tid, a, twB.data()); }
        bool third = false;
        for (u32 tid = 0; tid < 256; tid++) third = n2_second_round<MODE>(p, WORK.data(), tid, twB.data(), out_mul);
        if (third) for (u32 tid = 0; tid < 256; tid++) n2_third_round(p, WORK.data(), tid, out_mul);
        for (u32 r = 0; r < 256; r++) memcpy(out + n2_row_global(p, pos, r), &WORK[r * N2_ROW_WORDS], 128);
    }
}
template <int T> static int check_pow2() {
    const u64 c = gl_pow(2, T);
    const u64 edge[] = {0, 1, GL_P - 1, GL_P - 2, 0xFFFFFFFFULL, 0x100000000ULL, 0xFFFFFFFF00000000ULL, 0xFFFFFFFEFFFFFFFFULL, 0x8000000000000000ULL};
    for (u64 x : edge) if (gl_mul_pow2<T>(x) != gl_mul(x, c)) { printf("gl_mul_pow2<%d> mismatch on %llx\n", T, (unsigned long long)x); return 1; }
    for (int i = 0; i < 200000; i++) { u64 x = rnd(); if (gl_mul_pow2<T>(x) != gl_mul(x, c)) { printf("gl_mul_pow2<%d> mismatch on %llx\n", T, (unsigned long long)x); return 1; } }
    return 0;
}
int main() {
    srand(5);
    if (check_pow2<12>() || check_pow2<24>() || check_pow2<36>() || check_pow2<48>() || check_pow2<60>() || check_pow2<72>() || check_pow2<84>() ||
        check_pow2<4>() || check_pow2<31>() || check_pow2<33>() || check_pow2<63>() || check_pow2<65>() || check_pow2<95>()) return 1;
    printf("gl_mul_pow2 == gl_mul by 2^T for 13 shifts\n");
    std::vector<u64> twf = make_tw(20, false), twi = make_tw(20, true);
    for (int L = 16; L <= 20; L++) {
        const size_t n = size_t(1) << L;
        N2Pass p1 = {L, 0, 8, 1, 0, L - 8}, p2 = {L, 8, L - 8, 0, 0, 0};
        for (int inv = 0; inv < 2; inv++) {
            p1.inverse = p2.inverse = inv;
            std::vector<u64> x(n), y(n), ref;
            for (auto &v : x) v = rnd();
            x[0] = GL_P - 1; x[1] = 0; x[n - 1] = GL_P - 1;
            ref = x; ref_ntt(ref, L, inv);
            const u64 mul = inv ? gl_inv((u64)n % GL_P) : 1;
            run_pass<0>(p1, x.data(), y.data(), inv ? twi.data() : twf.data(), nullptr, 1);
            run_pass<2>(p2, y.data(), y.data(), inv ? twi.data() : twf.data(), nullptr, mul);
            for (size_t i = 0; i < n; i++) if (y[bitrev32((u32)i, L)] != gl_mul(ref[i], mul)) { printf("MISMATCH L=%d inv=%d i=%zu\n", L, inv, i); return 1; }
        }
        p1.inverse = p2.inverse = 0;
        if (L <= 18) {      // coset LDE: coefficients in bit-reversed storage, scale table in the same order
            std::vector<u64> coef(n), stor(n), scale(n), y(n), ref(n);
            for (auto &v : coef) v = rnd();
            const u64 c = gl_mul(GL_GEN, gl_pow(gl_root_of_unity(L + 3), 5));        // coset 5 of 8: g * w_{8n}^5
            { u64 s = 1; for (size_t i = 0; i < n; i++) { size_t q = bitrev32((u32)i, L); stor[q] = coef[i]; scale[q] = s; ref[i] = gl_mul(coef[i], s); s = gl_mul(s, c); } }
            ref_ntt(ref, L, false);
            N2Pass pb = p1; pb.in_bitrev = 1;
            run_pass<1>(pb, stor.data(), y.data(), twf.data(), scale.data(), 1);
            run_pass<2>(p2, y.data(), y.data(), twf.data(), nullptr, 1);
            for (size_t i = 0; i < n; i++) if (y[bitrev32((u32)i, L)] != ref[i]) { printf("LDE MISMATCH L=%d i=%zu\n", L, i); return 1; }
            // and without a scale table (plain transform from bit-reversed storage)
            ref = coef; ref_ntt(ref, L, false);
            run_pass<1>(pb, stor.data(), y.data(), twf.data(), nullptr, 1);
            run_pass<2>(p2, y.data(), y.data(), twf.data(), nullptr, 1);
            for (size_t i = 0; i < n; i++) if (y[bitrev32((u32)i, L)] != ref[i]) { printf("BITREV-IN MISMATCH L=%d i=%zu\n", L, i); return 1; }
        }
        printf("L=%d ok\n", L);
    }
    printf("ntt2 emulation == reference NTT for L = 16..20 (forward, inverse, coset LDE)\n");
}
