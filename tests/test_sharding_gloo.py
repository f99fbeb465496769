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

def test_shard_indices_partition():
    from b200_plonky2 import sharding
    for k in (0, 1, 62, 81, 100):
        for w in (1, 2, 4, 8):
            parts = [sharding.shard_indices(k, w, r) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(k))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
