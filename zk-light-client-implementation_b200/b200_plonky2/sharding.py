"""Multi-GPU sharding of the independent signature sub-proofs: signature index -> rank round-robin, one
all-gather of fixed-size digest records before the sequential recursive fold.  Replaces the reference's NATS fan-out
(near_bft_finality/src/prove_block_data/signatures.rs:144-274: PROVE_SIGNATURE / PROCESS_SIGNATURE_RESULT).
Backend-agnostic (nccl on GPUs, gloo in the CPU tests)."""
import numpy as np

RECORD_WORDS = 5        # [signature_index, 4 x u63 digest words]

def shard_indices(num_signatures, world_size, rank):
    """indices of the signatures proved by `rank` (round-robin keeps validator order inside every shard)."""
    return list(range(rank, num_signatures, world_size))

def digest_record(sig_index, proof_bytes):
    """fixed-size record of one sub-proof: index + the first wires-cap hash (4 field elements; int64-safe >> 1)."""
    words = [int.from_bytes(proof_bytes[8 * i:8 * i + 8], "little") >> 1 for i in range(4)]
    return [int(sig_index)] + words

def gather_records(records, dist, device, max_per_rank):
    """One all-gather of [max_per_rank, 5] int64 per rank; returns the records of all ranks sorted by signature index
    (the fold must consume them in validator order: signatures.rs:97-126)."""
    import torch
    t = torch.full((max_per_rank, RECORD_WORDS), -1, dtype=torch.int64, device=device)
    if records:
        t[:len(records)] = torch.tensor(records, dtype=torch.int64, device=device)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        allr = [t]
    else:
        allr = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(allr, t)
    rows = torch.cat(allr).cpu().numpy()
    rows = rows[rows[:, 0] >= 0]
    return rows[np.argsort(rows[:, 0], kind="stable")]
