"""N > 1 host logic on CPU: world_size-2 gloo run of the signature sharding + the single all-gather."""
import os, socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p

def _worker(rank, world, port, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200_plonky2 import sharding
    mine = sharding.shard_indices(k, world, rank)
    recs = [sharding.digest_record(i, bytes([(i * 7 + j) % 256 for j in range(64)])) for i in mine]
    rows = sharding.gather_records(recs, dist, "cpu", -(-k // world))
    np.save(os.path.join(out_dir, f"rows{rank}.npy"), rows)
    dist.destroy_process_group()

@pytest.mark.parametrize("k", [73, 100, 1])
def test_shard_and_gather_world2(tmp_path, k):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = os.pathsep.join([root, os.path.join(root, "zk-light-client-implementation_b200"), os.environ.get("PYTHONPATH", "")])
    port = _free_port()
    mp.spawn(_worker, args=(2, port, k, str(tmp_path)), nprocs=2, join=True)
    from b200_plonky2 import sharding
    r0 = np.load(tmp_path / "rows0.npy"); r1 = np.load(tmp_path / "rows1.npy")
    assert np.array_equal(r0, r1)                         # every rank ends with the same ordered record list
    assert list(r0[:, 0]) == list(range(k))               # validator order restored, nothing lost or duplicated
    for i in range(k):
        assert list(r0[i]) == sharding.digest_record(i, bytes([(i * 7 + j) % 256 for j in range(64)]))

def _worker_proofs(rank, world, port, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200_plonky2 import sharding
    mine = sharding.shard_indices(k, world, rank)
    proofs = {i: bytes([(i * 13 + j) % 251 for j in range(1000 + 37 * i)]) for i in mine}      # ragged lengths
    got = sharding.gather_proofs(proofs, dist, "cpu", -(-k // world), 1000 + 37 * k)
    import pickle
    pickle.dump(got, open(os.path.join(out_dir, f"proofs{rank}.pkl"), "wb"))
    dist.destroy_process_group()

@pytest.mark.parametrize("k", [7, 1])
def test_gather_full_proofs_world2(tmp_path, k):
    """the real payload: complete (ragged-length) proof bytes of every sub-proof reach every rank in one all-gather"""
    import pickle
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = os.pathsep.join([root, os.path.join(root, "zk-light-client-implementation_b200"), os.environ.get("PYTHONPATH", "")])
    mp.spawn(_worker_proofs, args=(2, _free_port(), k, str(tmp_path)), nprocs=2, join=True)
    g0 = pickle.load(open(tmp_path / "proofs0.pkl", "rb")); g1 = pickle.load(open(tmp_path / "proofs1.pkl", "rb"))
    assert g0 == g1 and sorted(g0) == list(range(k))
    for i in range(k): assert g0[i] == bytes([(i * 13 + j) % 251 for j in range(1000 + 37 * i)])

def test_shard_indices_partition():
    from b200_plonky2 import sharding
    for k in (0, 1, 62, 81, 100):
        for w in (1, 2, 4, 8):
            parts = [sharding.shard_indices(k, w, r) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(k))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
