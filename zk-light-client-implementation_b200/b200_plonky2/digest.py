"""circuit_digest = H(constants_sigmas_cap || H_pad(domain_separator = []) || degree_bits), PoseidonGoldilocks hasher
(upstream plonky2 CircuitBuilder::build, not vendored).  Host-side, tiny."""
import numpy as np
from . import poseidon_np

def hash_pad(inputs):
    x = list(inputs) + [1]
    while (len(x) + 1) % 8 != 0: x.append(0)
    x.append(1)
    return poseidon_np.hash_no_pad(x)

def circuit_digest(cap, degree_bits):
    parts = [int(v) for v in np.asarray(cap).reshape(-1)] + hash_pad([]) + [degree_bits]
    return poseidon_np.hash_no_pad(parts)
