"""Static task graph of one BFT-finality block proof, as the reference's orchestration builds it
(near_bft_finality/src/prove_bft/bft.rs:38-565 prove_block_bft; block_finality.rs:200-731; prove_block_data/
signatures.rs:43-141 prove_approvals; header_bphash.rs:34-139 prove_block_header / bp_hash; keys_stakes.rs:18-266;
primitives.rs:32,126,176), reduced to what the prover sees: WHICH proof shapes are produced, in which dependency order.

  stage S  k signature sub-proofs (ed25519-shaped, 2^17 x 234)                      independent  -> sharded over all ranks
  stage A  the other leaf proofs of the block, all independent of the signatures:
             7 block headers x 4 SHA-256 sub-proofs (inner_lite 208 B, inner_rest, hash-of-hashes, with prev_hash:
               header_bphash.rs:44-93 -> sha256_proof_u32, 2..5 blocks each)        28 x sha-shaped 2^12
             next_bp_hash = SHA-256 over the validator list (~110 blocks, header_bphash.rs:125-139)   3 x sha-shaped 2^15
             prove_eq_array / prove_consecutive_heights / two_thirds (primitives.rs)  9 x tiny 2^5
             stake sum over the validator list (keys_stakes.rs:18)                   1 x rec-shaped 2^12
  stage C  the signature chain: k recursive_proof steps, strictly sequential (signatures.rs:97-105)   rank 0
  stage F  the remaining 55 recursive aggregations (headers, heights, bp_hash, stake into the block proof; bft.rs) and the
           PoseidonBN128 wrap (src/bin/prove_block.rs:279-300)                       rank 0, after stage A's proofs arrived

Stage A does not depend on stage S or C, so with N > 1 ranks it runs on ranks 1..N-1 WHILE rank 0 walks the chain; with one
rank it runs on a second proving context of the same GPU, concurrently with the chain.  Two collectives move proof bytes:
after S (sub-proofs to rank 0 for the chain) and after A.  This module only describes and schedules; the proofs are made by
b200_prove through CircuitData.prove.
"""
from . import sharding

LEAF_SHAPES = [("sha", 12)] * 28 + [("sha", 15)] * 3 + [("tiny_pi", 5)] * 9 + [("rec", 12)]
FINAL_AGGREGATIONS = 55

def leaf_tasks():
    """[(task_id, kind, degree_bits)] of stage A"""
    return [(i, kind, db) for i, (kind, db) in enumerate(LEAF_SHAPES)]

def assign_leaf_tasks(world, rank):
    """stage A tasks of `rank`: with one rank everything is local; otherwise rank 0 is busy with the chain and the tasks go
    round-robin (heaviest first) to ranks 1..N-1"""
    tasks = sorted(leaf_tasks(), key=lambda t: -t[2])
    if world == 1: return tasks
    return [t for j, t in enumerate(tasks) if 1 + j % (world - 1) == rank]

def max_leaf_tasks_per_rank(world):
    return max(len(assign_leaf_tasks(world, r)) for r in range(world))

def summary(k):
    return {"signature_proofs": k, "leaf_proofs": len(LEAF_SHAPES), "chain_recursions": k, "final_aggregations": FINAL_AGGREGATIONS,
            "wrap_proofs": 1, "leaf_shapes": {"sha 2^12": 28, "sha 2^15": 3, "tiny 2^5": 9, "stake (rec-shaped) 2^12": 1}}
