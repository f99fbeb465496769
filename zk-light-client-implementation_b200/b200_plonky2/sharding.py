"""Multi-GPU sharding of the independent signature sub-proofs: signature index -> rank round-robin, one
all-gather of the sub-proofs before the sequential recursive fold.  Replaces the reference's NATS fan-out
(near_bft_finality/src/prove_block_data/signatures.rs:144-274: PROVE_SIGNATURE / PROCESS_SIGNATURE_RESULT, whose result
message carries the whole serialized proof + verifier data: signatures.rs:225-232).
Backend-agnostic (nccl on GPUs, gloo in the CPU tests).

Two payloads: gather_proofs moves the COMPLETE proof bytes of every sub-proof to every rank in one collective (what
recursive_proof needs as witness: signatures.rs:97-105); gather_records moves only (index, first cap digest) records (the
digest-only variant north_star describes)."""
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


# ---- full sub-proof payload ----
def pack_proofs(proofs, max_count, stride, device, pin=False):
    """proofs: {index: bytes} of this rank -> uint8 tensor [max_count, 16 + stride]: per row
    int64 index (-1 = empty), int64 length, proof bytes (zero padded)."""
    import torch
    stride = (stride + 7) & ~7                                   # rows stay 8-byte aligned (the header is viewed as int64)
    t = torch.zeros((max_count, 16 + stride), dtype=torch.uint8, pin_memory=pin)
    hdr = t[:, :16].view(torch.int64)
    hdr[:, 0] = -1
    for row, (idx, b) in enumerate(sorted(proofs.items())):
        if len(b) > stride: raise ValueError(f"proof {idx}: {len(b)} bytes > stride {stride}")
        hdr[row, 0] = idx; hdr[row, 1] = len(b)
        t[row, 16:16 + len(b)] = torch.frombuffer(bytearray(b), dtype=torch.uint8)
    return t.to(device, non_blocking=True) if str(device) != "cpu" else t

def gather_proofs(proofs, dist, device, max_count, stride, group=None, group_size=None):
    """ONE all-gather of every rank's packed proofs; returns {index: bytes} of all ranks (identical on every rank).
    group / group_size: a torch.distributed sub-group (and its size) for gathers among part of the ranks."""
    import torch
    t = pack_proofs(proofs, max_count, stride, device)
    n = group_size if group is not None else (dist.get_world_size() if dist is not None and dist.is_initialized() else 1)
    if dist is None or not dist.is_initialized() or n == 1:
        allt = t.unsqueeze(0)
    else:
        allt = torch.empty((n * t.shape[0], t.shape[1]), dtype=torch.uint8, device=t.device)
        dist.all_gather_into_tensor(allt, t, group=group)        # member r's rows land at [r * max_count, (r + 1) * max_count)
        allt = allt.view(n, t.shape[0], t.shape[1])
    h = allt.cpu()
    out = {}
    for r in range(h.shape[0]):
        hdr = h[r][:, :16].contiguous().view(torch.int64)
        for row in range(h.shape[1]):
            idx, ln = int(hdr[row, 0]), int(hdr[row, 1])
            if idx >= 0: out[idx] = bytes(h[r][row, 16:16 + ln].numpy().tobytes())
    return out

def payload_bytes(max_count, stride, world):
    return world * max_count * (16 + ((stride + 7) & ~7))
