"""BN254 scalar-field (F_r) NTT with Python integers: the oracle for b200_fr_ntt (next row, SURVEY §8f-3: the FFTs of
groth16.Prove computing the quotient H; gnark-crypto v0.12.2 ecc/bn254/fr/fft is not vendored -- its published domain
parameters are restated here: multiplicative generator 5, 2-adicity 28, the 2^28-th root of unity below).
TEST INFRASTRUCTURE.  Natural order in and out.
  forward:          A[k] = sum_i a[i] w^(ik)                    (fft.Domain.FFT, DIF + bit reversal)
  forward, coset:   A[k] = sum_i a[i] (g w^k)^i                 (fft.OnCoset(): a[i] *= g^i first)
  inverse:          a[i] = n^-1 sum_k A[k] w^(-ik)              (FFTInverse), on the coset additionally a[i] *= g^-i
"""
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
GEN = 5
TWO_ADICITY = 28
ROOT_2_28 = 19103219067921713944291392827692070036145651957329286315305642004821462161904

def root_of_unity(log_n):
    assert 0 <= log_n <= TWO_ADICITY
    return pow(ROOT_2_28, 1 << (TWO_ADICITY - log_n), R)

def _bitrev(i, bits):
    return int(format(i, "0%db" % bits)[::-1], 2) if bits else 0

def ntt(a, inverse=False, coset=False):
    n = len(a); log_n = n.bit_length() - 1
    assert n == 1 << log_n
    a = [x % R for x in a]
    if coset and not inverse:
        g = 1
        for i in range(n): a[i] = a[i] * g % R; g = g * GEN % R
    w = root_of_unity(log_n)
    if inverse: w = pow(w, -1, R)
    # iterative Cooley-Tukey (decimation in time) on a bit-reversed copy
    out = [a[_bitrev(i, log_n)] for i in range(n)]
    m = 1
    while m < n:
        wm = pow(w, n // (2 * m), R)
        for k in range(0, n, 2 * m):
            t = 1
            for j in range(m):
                u, v = out[k + j], out[k + j + m] * t % R
                out[k + j] = (u + v) % R; out[k + j + m] = (u - v) % R
                t = t * wm % R
        m *= 2
    if inverse:
        ninv = pow(n, -1, R)
        out = [x * ninv % R for x in out]
        if coset:
            gi = pow(GEN, -1, R); g = 1
            for i in range(n): out[i] = out[i] * g % R; g = g * gi % R
    return out

def naive_dft(a, coset=False):
    n = len(a); log_n = n.bit_length() - 1
    w = root_of_unity(log_n); g = GEN if coset else 1
    return [sum(a[i] * pow(g * pow(w, k, R) % R, i, R) for i in range(n)) % R for k in range(n)]
