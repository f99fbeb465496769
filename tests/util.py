import numpy as np
P = (1 << 64) - (1 << 32) + 1

def rand_field(shape, seed):
    """uniform-ish canonical field elements, seeded (splitmix-style via numpy)."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 63, size=shape, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=shape, dtype=np.uint64)
    a = np.where(a >= np.uint64(P), a - np.uint64(P), a)
    return a.astype(np.uint64)

def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
