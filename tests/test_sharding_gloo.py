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

# ---- the block task graph (b200_plonky2/block_dag.py): stage A lanes on ranks 1..N-1, gathered in a sub-group, joined on rank 1 ----
def test_block_dag_partitions_every_task_once():
    from b200_plonky2 import block_dag as bd
    assert bd.HEADER_AGGREGATIONS + bd.JOIN_AGGREGATIONS + bd.POST_CHAIN_FOLDS == bd.FINAL_AGGREGATIONS
    lane_tasks = [t for _, tasks in bd.lanes() for t in tasks]
    assert [t[0] for t in lane_tasks] == list(range(bd.num_lane_tasks()))
    non_rec = sorted((k, d) for _, k, d in lane_tasks if not (k == "rec" and d == 13))
    assert non_rec == sorted(bd.LEAF_SHAPES)                                   # the 41 leaf proofs, each in exactly one lane
    assert sum(1 for _, k, d in lane_tasks if (k, d) == ("rec", 13)) == bd.HEADER_AGGREGATIONS
    assert [t[0] for t in bd.join_tasks()] == list(range(bd.num_lane_tasks(), bd.num_lane_tasks() + bd.JOIN_AGGREGATIONS))
    for w in (1, 2, 3, 4, 8):
        got = [l for r in range(w) for l, _ in bd.assign_lanes(w, r)]
        assert sorted(got) == list(range(len(bd.lanes())))                     # every lane on exactly one rank
        if w > 1: assert bd.assign_lanes(w, 0) == []                           # rank 0 walks the chain
        assert bd.join_rank(w) in bd.stage_a_ranks(w)
        per_rank = [sum(len(t) for _, t in bd.assign_lanes(w, r)) + (bd.JOIN_AGGREGATIONS if r == bd.join_rank(w) else 0) for r in range(w)]
        assert max(per_rank) == bd.max_stage_a_tasks_per_rank(w) and sum(per_rank) == bd.num_lane_tasks() + bd.JOIN_AGGREGATIONS

def _worker_stage_a(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200_plonky2 import sharding, block_dag as bd
    a_ranks = bd.stage_a_ranks(world)
    sub = dist.new_group(ranks=a_ranks)                                        # every rank calls it
    fake = lambda tid: bytes([(tid * 11 + j) % 253 for j in range(300 + 5 * tid)])
    a_out = {}
    if rank in a_ranks:
        for _, tasks in bd.assign_lanes(world, rank):
            for tid, _, _ in tasks: a_out[tid] = fake(tid)
        lanes_all = sharding.gather_proofs(dict(a_out), dist, "cpu", bd.max_lane_tasks_per_rank(world), 1000, group=sub, group_size=len(a_ranks))
        assert sorted(lanes_all) == list(range(bd.num_lane_tasks()))           # the joining rank sees every lane result
        if rank == bd.join_rank(world):
            for tid, _, _ in bd.join_tasks(): a_out[tid] = fake(tid)
    all_a = sharding.gather_proofs(a_out, dist, "cpu", bd.max_stage_a_tasks_per_rank(world), 1000)
    import pickle
    pickle.dump(all_a, open(os.path.join(out_dir, f"a{rank}.pkl"), "wb"))
    dist.destroy_process_group()

def test_stage_a_subgroup_gather_world3(tmp_path):
    """rank 0 stays out of the lane gather (it is walking the chain); ranks 1..2 exchange lane results in a sub-group, rank 1
    adds the joins, and the closing all-gather hands every stage-A proof to every rank"""
    import pickle
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = os.pathsep.join([root, os.path.join(root, "zk-light-client-implementation_b200"), os.environ.get("PYTHONPATH", "")])
    mp.spawn(_worker_stage_a, args=(3, _free_port(), str(tmp_path)), nprocs=3, join=True)
    from b200_plonky2 import block_dag as bd
    got = [pickle.load(open(tmp_path / f"a{r}.pkl", "rb")) for r in range(3)]
    assert got[0] == got[1] == got[2]
    assert sorted(got[0]) == list(range(bd.num_lane_tasks() + bd.JOIN_AGGREGATIONS))
    for tid, b in got[0].items(): assert b == bytes([(tid * 11 + j) % 253 for j in range(300 + 5 * tid)])
