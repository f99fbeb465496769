// Host run of the limb-resident Poseidon permutation of poseidon.cuh (poseidon_permute_limb: the code the Merkle kernels
// execute, compiled for the host: same limb arithmetic, the PTX carry chains replaced by their C equivalents) against the
// textbook poseidon_permute, with limb-range assertions.  Built and run by tests/test_host_emulations.py (no device needed).
#include "poseidon.cuh"
#include <cstdio>
#include <cstdlib>
static u64 rnd() { return ((u64)rand() << 42) ^ ((u64)rand() << 21) ^ (u64)rand(); }
int main() {
    srand(7);
    const u64 edge[] = {0, 1, GL_P - 1, GL_P - 2, 0xFFFFFFFFULL, 0xFFFFFFFF00000000ULL, 0x100000000ULL, 0x1FFFFFULL, 0x3FFFFFFFFFFULL, 0x8000000000000000ULL};
    const int NE = sizeof edge / sizeof edge[0];
    for (int t = 0; t < 20000; t++) {
        u64 a[12], b[12];
        for (int i = 0; i < 12; i++) {
            u64 v = rnd() % GL_P;
            if (t < NE) v = edge[t];
            else if (t < 4 * NE) v = edge[(t + i * (t / NE)) % NE];
            a[i] = b[i] = v;
        }
        poseidon_permute(a); poseidon_permute_limb(b, 1);
        for (int i = 0; i < 12; i++) if (a[i] != b[i]) { printf("MISMATCH t=%d i=%d %llx %llx\n", t, i, (unsigned long long)a[i], (unsigned long long)b[i]); return 1; }
    }
    // single pieces on extreme limb values
    for (int t = 0; t < 200000; t++) {
        u32 y0 = (u32)(rnd() % (264u << 22)), y1 = (u32)(rnd() % (264u << 22)), y2 = (u32)(rnd() % (265u << 22));
        if (t == 0) { y0 = y1 = (264u << 22) - 1; y2 = (265u << 22) - 1; }
        if (t == 1) { y0 = y1 = y2 = 0; }
        unsigned __int128 v = (unsigned __int128)y0 + ((unsigned __int128)y1 << 21) + ((unsigned __int128)y2 << 42) + POSEIDON_L_HB;
        u64 want = (u64)(v % GL_P);
        if (gl_canon(pl_recombine(y0, y1, y2)) != want) { printf("recombine mismatch\n"); return 1; }
        i32 a = (i32)y0, b = (i32)y1, c = (i32)y2; pl_normalise(a, b, c);
        if (a < 0 || b < 0 || c < 0 || a >= (1 << 21) + POSEIDON_L_HB + 1 || b >= (1 << 21) + (1 << 20) || c >= (1 << 22)) { printf("normalise range\n"); return 1; }
        unsigned __int128 w = (unsigned __int128)(u32)a + ((unsigned __int128)(u32)b << 21) + ((unsigned __int128)(u32)c << 42);
        if ((u64)(w % GL_P) != want) { printf("normalise value\n"); return 1; }
    }
    printf("limb-resident permutation == poseidon_permute on 20000 states\n");
}
