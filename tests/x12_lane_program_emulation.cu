// Host emulation of the 12-lane cooperative Poseidon round program of merkle.cu (poseidon_permute_x12: lane w holds state
// word w, the MDS row is gathered from the other lanes by shuffles, here array reads) against poseidon_permute.
// Built and run by tests/test_host_emulations.py with nvcc in host-only mode (no device needed).
#include "poseidon.cuh"
#include <cstdio>
#include <cstdlib>
static void permute_x12_emulated(u64 s[12]) {
    const u32 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    u64 x[12]; for (int w = 0; w < 12; w++) x[w] = s[w];
    for (int r = 0; r < 30; r++) {
        const bool full = (r < 4) || (r >= 26);
        for (int w = 0; w < 12; w++) {          // per-lane: constant + S-box
            u64 c = h_POSEIDON_RC[12 * r + w], v = x[w] + c;
            if (v < c) v += GL_EPS;
            if (full || w == 0) v = poseidon_sbox(v);
            x[w] = v;
        }
        u64 y[12];
        for (int w = 0; w < 12; w++) {          // per-lane: MDS row through "shuffles"
            u64 al = 0, ah = 0; int src = w;
            for (int i = 0; i < 12; i++) {
                u64 nb = x[src];                // = shfl(x, base + src)
                al += (u64)(u32)nb * C[i]; ah += (nb >> 32) * C[i];
                if (++src == 12) src = 0;
            }
            if (w == 0) { al += (u64)(u32)x[0] * 8u; ah += (x[0] >> 32) * 8u; }
            y[w] = poseidon_reduce96(al, ah);
        }
        for (int w = 0; w < 12; w++) x[w] = y[w];
    }
    for (int w = 0; w < 12; w++) s[w] = gl_canon(x[w]);
}
int main() {
    srand(1);
    for (int t = 0; t < 2000; t++) {
        u64 a[12], b[12];
        for (int i = 0; i < 12; i++) { u64 v = ((u64)rand() << 42) ^ ((u64)rand() << 21) ^ rand(); if (t < 3) v = t == 0 ? 0 : (t == 1 ? GL_P - 1 : 0xFFFFFFFFULL); a[i] = b[i] = v % GL_P; }
        poseidon_permute(a); permute_x12_emulated(b);
        for (int i = 0; i < 12; i++) if (a[i] != b[i]) { printf("MISMATCH t=%d i=%d\n", t, i); return 1; }
    }
    printf("x12 emulation == poseidon_permute on 2000 states\n");
}
