// Index math and register butterflies of the bulk-copy (TMA) staged NTT passes (ntt2.cu).  Host + device: the kernel runs
// these with shared-memory buffers and its thread index, tests/ntt2_emulation.cu runs the same functions over plain arrays
// for tid = 0..255 round by round, so the thread <-> element mapping, the twiddle indexing and the row layouts are checked
// on the CPU; only the asynchronous copy / mbarrier plumbing is device-only.
//
// A pass covers B consecutive DIF layers of a 2^L-point transform on one tile of 4096 elements:
//   strided pass   (not the last; B = 8): tile = 256 `mid` values x 16 consecutive `low` values, element (mid, c) is the
//                   natural index top * 2^(L - l_begin) + mid * S + low0 + c, S = 2^(L - l_begin - 8)
//   contiguous pass (the last; 8 <= B <= 12): tile = 4096 consecutive elements = 4096 / 2^B sub-transforms of 2^B points
// Tile-local index idx (12 bits): strided idx = mid * 16 + c; contiguous idx = offset inside the tile.  Shared-memory layout
// of a buffer: rows of 16 elements, 144 bytes apart (16-byte pad): word offset = (idx >> 4) * 18 + (idx & 15).  With this
// stride (a) every row is a legal 16-byte-aligned bulk-copy destination, (b) 64-bit accesses of a warp that walk 16
// consecutive columns of 2 rows and (c) 128-bit accesses of a warp that walk 32 consecutive rows both hit the minimum
// number of shared-memory wavefronts.
// A pass is 2 or 3 ROUNDS; in a round every thread pulls 2^R elements (one radix-2^R group) into registers, runs R
// butterfly layers on them and writes them back.  Rounds act on idx bits [q, q + R): (q, R) = (8, B - 8) if B > 8, then
// (4, 4), then (0, 4).
#pragma once
#include "gl.cuh"

#define N2_THREADS 256
#define N2_TILE 4096
#define N2_ROW_WORDS 18
#define N2_BUF_WORDS (256 * N2_ROW_WORDS)              /* 36 864 bytes */
#define N2_OFF(idx) ((((idx) >> 4) * N2_ROW_WORDS) + ((idx) & 15))

struct N2Pass {
    int L, l_begin, B;              // layers [l_begin, l_begin + B)
    int strided;                    // 1: strided pass (B == 8); 0: contiguous last pass
    int in_bitrev;                  // strided first pass of an LDE: input in bit-reversed storage (l_begin == 0)
    int logS;                       // L - l_begin - B  (0 for the contiguous pass)
    int inverse;                    // inverse transform (twiddle table T.inv): selects the shift table of the multiplier-free last round
};

GL_HD u32 n2_rev4(u32 x) { return ((x & 1) << 3) | ((x & 2) << 1) | ((x & 4) >> 1) | ((x & 8) >> 3); }

// ---- global <-> tile row mapping: row r (0..255) of tile `tile` of an item = 16 consecutive u64 ----
// returns the element offset (inside the item) of the first element of the row as it sits in GLOBAL memory
// natural layouts (input of an ordinary pass, output of every pass)
GL_HD u64 n2_row_global(const N2Pass &p, u32 tile, u32 r) {
    if (!p.strided) return (u64)tile * N2_TILE + (u64)r * 16;
    const u64 S = u64(1) << p.logS;
    const u64 lows = S >> 4;                                    // tiles per `top`
    const u64 top = tile / lows, low0 = (tile % lows) << 4;
    return (top << (p.L - p.l_begin)) + (u64)r * S + low0;     // row = mid
}
// bit-reversed storage (first pass of an LDE, l_begin == 0): row r = ll * 16 + ch holds, for low = low0 + ll and
// mid & 15 == rev4(ch), the 16 values mid >> 4 == rev4(k), k = 0..15  (one radix-16 group of the first round, bit-reversed)
GL_HD u64 n2_row_global_bitrev(const N2Pass &p, u32 tile, u32 r) {
    const u32 ll = r >> 4, ch = r & 15;
    const int lowbits = p.L - 8;                                // low = tile * 16 + ll
    const u32 low = tile * 16 + ll;
    return ((u64)bitrev32(low, lowbits) << 8) + ch * 16;
}

// ---- register butterflies: R DIF layers on 2^R values; tws[(1 << s) - 1 + k], k < 2^s, is the twiddle of sub-layer s ----
template <int R>
GL_HD void n2_radix(u64 (&x)[1 << R], const u64 *tws) {
#pragma unroll
    for (int s = R - 1; s >= 0; s--) {
#pragma unroll
        for (int e = 0; e < (1 << R); e++) {
            if (e & (1 << s)) continue;
            const u64 w = tws[(1 << s) - 1 + (e & ((1 << s) - 1))];
            const u64 u = x[e], v = x[e | (1 << s)];
            x[e] = gl_add(u, v);
            x[e | (1 << s)] = gl_mul(gl_sub(u, v), w);
        }
    }
}
// The round on idx bits [0, 4) of the contiguous pass is a plain 16-point DFT: its twiddles are the 16th roots of unity
// w_16^k = 2^(156 k) (inverse transform: 2^(36 k)), i.e. shifts (gl_mul_pow2) and sign flips -- 17 shift-multiplies instead
// of 32 full products per 16 elements.  Same values as n2_radix<4> with the table twiddles tw[2^s + k].
template <int INV>
GL_HD void n2_radix16_pow2(u64 (&x)[16]) {
#pragma unroll
    for (int s = 3; s >= 0; s--) {
#pragma unroll
        for (int e = 0; e < 16; e++) {
            if (e & (1 << s)) continue;
            const int k = e & ((1 << s) - 1);
            const int E = ((INV ? 36 : 156) * k * (8 >> s)) % 192;          // w_{2^(s+1)}^k = w_16^(k 2^(3-s)) = 2^E
            const u64 u = x[e], v = x[e | (1 << s)];
            x[e] = gl_add(u, v);
            const u64 d = E >= 96 ? gl_sub(v, u) : gl_sub(u, v);
            u64 r;
            switch (E % 96) {
            case 0: r = d; break;
            case 12: r = gl_mul_pow2<12>(d); break;
            case 24: r = gl_mul_pow2<24>(d); break;
            case 36: r = gl_mul_pow2<36>(d); break;
            case 48: r = gl_mul_pow2<48>(d); break;
            case 60: r = gl_mul_pow2<60>(d); break;
            case 72: r = gl_mul_pow2<72>(d); break;
            default: r = gl_mul_pow2<84>(d); break;
            }
            x[e | (1 << s)] = r;
        }
    }
}
// twiddle-table index of sub-layer s of a round on idx bits [q, q + R) for a group whose elements below bit q read `jbase`
// (jbase = natural index of the group's first element modulo 2^(P + q), P = logS): tw[2^beta + jbase + k * 2^(P + q)], beta = P + q + s
GL_HD u64 n2_tw_index(int P, int q, int s, u64 jbase, u32 k) {
    return (u64(1) << (P + q + s)) + jbase + ((u64)k << (P + q));
}
template <int R>
GL_HD void n2_load_tws(const u64 *tw, int P, int q, u64 jbase, u64 (&tws)[(1 << R) - 1]) {
#pragma unroll
    for (int s = 0; s < R; s++)
#pragma unroll
        for (int k = 0; k < (1 << s); k++) tws[(1 << s) - 1 + k] = tw[n2_tw_index(P, q, s, jbase, k)];
}

// ---- one thread's share of a round.  in / outb: buffers in the padded row layout (may alias: a thread only touches its own
// group).  Rounds on bits [q, q + R), q >= 4: thread task u (12 - R bits): idx = ((u >> q) << (q + R)) | (e << q) | (u & (2^q - 1)).
template <int R>
GL_HD void n2_round_gather(const u64 *in, u64 *outb, u32 task, int q, const u64 *tws, u64 out_mul) {
    const u32 lo = task & ((1u << q) - 1), hi = task >> q;
    const u32 base = (hi << (q + R)) | lo;
    u64 x[1 << R];
#pragma unroll
    for (int e = 0; e < (1 << R); e++) x[e] = in[N2_OFF(base | ((u32)e << q))];
    n2_radix<R>(x, tws);
    if (out_mul != 1) {
#pragma unroll
        for (int e = 0; e < (1 << R); e++) x[e] = gl_mul(x[e], out_mul);
    }
#pragma unroll
    for (int e = 0; e < (1 << R); e++) outb[N2_OFF(base | ((u32)e << q))] = x[e];
}
// last round (bits [0, 4)): thread = row; 128-bit accesses; multiplier-free 16-point DFT
GL_HD void n2_round_row(const u64 *in, u64 *outb, u32 row, int inverse, u64 out_mul) {
    u64 x[16];
    const u64 *src = in + row * N2_ROW_WORDS;
#pragma unroll
    for (int e = 0; e < 16; e++) x[e] = src[e];
    if (inverse) n2_radix16_pow2<1>(x); else n2_radix16_pow2<0>(x);
    if (out_mul != 1) {
#pragma unroll
        for (int e = 0; e < 16; e++) x[e] = gl_mul(x[e], out_mul);
    }
    u64 *dst = outb + row * N2_ROW_WORDS;
#pragma unroll
    for (int e = 0; e < 16; e++) dst[e] = x[e];
}
// first round of the bit-reversed (LDE) pass: thread = row r of the chunk layout (n2_row_global_bitrev): x_e = row[rev4(e)]
// (times the coset scale: sc_row = the 16 scale factors of this row, read straight from the table in global memory -- same
// storage order as the coefficients), butterflies on mid bits [4, 8), results to the standard layout (mid * 16 + c)
GL_HD void n2_round_bitrev_first(const u64 *in, const u64 *sc_row, u64 *outb, u32 r, const u64 *tws) {
    const u32 c = r >> 4, m_lo = n2_rev4(r & 15);
    const u64 *src = in + r * N2_ROW_WORDS;
    u64 raw[16], x[16];
    if (sc_row) {
        u64 sv[16];
#ifdef __CUDA_ARCH__
#pragma unroll
        for (int k = 0; k < 8; k++) { const ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2 *>(sc_row) + k); sv[2 * k] = t.x; sv[2 * k + 1] = t.y; }
#else
        for (int k = 0; k < 16; k++) sv[k] = sc_row[k];
#endif
#pragma unroll
        for (int k = 0; k < 16; k++) raw[k] = gl_mul(src[k], sv[k]);
    } else {
#pragma unroll
        for (int k = 0; k < 16; k++) raw[k] = src[k];
    }
#pragma unroll
    for (int e = 0; e < 16; e++) x[e] = raw[((e & 1) << 3) | ((e & 2) << 1) | ((e & 4) >> 1) | ((e & 8) >> 3)];
    n2_radix<4>(x, tws);
#pragma unroll
    for (int e = 0; e < 16; e++) outb[N2_OFF((((u32)e << 4) | m_lo) * 16 + c)] = x[e];
}

// ---- per-thread round drivers shared by the kernel (ntt2.cu) and the host emulation ----
// MODE 0: strided pass, natural input; 1: strided first pass of an LDE (bit-reversed input, optional coset scale); 2: contiguous last pass
// twiddles of the FIRST round of a unit (registers in the kernel); pos = tile position inside the item
template <int MODE>
GL_HD void n2_build_twsA(const N2Pass &p, const u64 *tw, u32 pos, u32 tid, u64 (&twsA)[15]) {
    if (MODE == 2) {
        if (p.B > 8) {                                           // round on idx bits [8, B): task & 255 == tid in every iteration
            const int RA = p.B - 8;
            for (int s = 0; s < RA; s++) for (int k = 0; k < (1 << s); k++) twsA[(1 << s) - 1 + k] = tw[n2_tw_index(0, 8, s, tid, k)];
        }
        return;
    }
    const u64 S = u64(1) << p.logS;
    const u64 low0 = ((u64)pos % (S >> 4)) << 4;
    const u32 c = MODE == 1 ? tid >> 4 : tid & 15, m_lo = MODE == 1 ? n2_rev4(tid & 15) : tid >> 4;
    n2_load_tws<4>(tw, p.logS - 4, 8, (u64)m_lo * S + low0 + c, twsA);     // mid bits [4, 8) = idx bits [8, 12)
}
// shared round table twB[c * 16 + e], c < 16, e < 15: the radix-16 round on idx bits [4, 8)
template <int MODE>
GL_HD u64 n2_twB_entry(const N2Pass &p, const u64 *tw, u32 pos, u32 c, u32 e) {
    int s = 0; while ((2u << s) <= e + 1) s++;                   // e + 1 in [2^s, 2^(s+1))
    const u32 k = e + 1 - (1u << s);
    if (MODE == 2) return tw[n2_tw_index(0, 4, s, c, k)];
    const u64 S = u64(1) << p.logS;
    const u64 low0 = ((u64)pos % (S >> 4)) << 4;
    return tw[n2_tw_index(p.logS - 4, 4, s, low0 + c, k)];       // mid bits [0, 4)
}
template <int MODE>
GL_HD void n2_first_round(const N2Pass &p, const u64 *IN, const u64 *sc_row, u64 *WORK, u32 tid, const u64 (&twsA)[15], const u64 *twB) {
    if (MODE == 1) { n2_round_bitrev_first(IN, sc_row, WORK, tid, twsA); return; }
    if (MODE == 0) { n2_round_gather<4>(IN, WORK, tid, 8, twsA, 1); return; }
    if (p.B > 8) {
        const int RA = p.B - 8, iters = 16 >> RA;               // 4096 / 2^RA tasks over 256 threads
        for (int it = 0; it < iters; it++) {
            const u32 task = tid + 256u * it;
            switch (RA) {
            case 1: n2_round_gather<1>(IN, WORK, task, 8, twsA, 1); break;
            case 2: n2_round_gather<2>(IN, WORK, task, 8, twsA, 1); break;
            case 3: n2_round_gather<3>(IN, WORK, task, 8, twsA, 1); break;
            default: n2_round_gather<4>(IN, WORK, task, 8, twsA, 1); break;
            }
        }
    } else n2_round_gather<4>(IN, WORK, tid, 4, twB + (tid & 15) * 16, 1);
}
// second round: in place on WORK.  Returns true if a third round (n2_third_round) follows after a barrier.
template <int MODE>
GL_HD bool n2_second_round(const N2Pass &p, u64 *WORK, u32 tid, const u64 *twB, u64 out_mul) {
    if (MODE != 2) { n2_round_gather<4>(WORK, WORK, tid, 4, twB + (tid & 15) * 16, out_mul); return false; }
    if (p.B > 8) { n2_round_gather<4>(WORK, WORK, tid, 4, twB + (tid & 15) * 16, 1); return true; }
    n2_round_row(WORK, WORK, tid, p.inverse, out_mul);
    return false;
}
GL_HD void n2_third_round(const N2Pass &p, u64 *WORK, u32 tid, u64 out_mul) { n2_round_row(WORK, WORK, tid, p.inverse, out_mul); }
