"""Static task graph of one BFT-finality block proof, as the reference's orchestration builds it
(near_bft_finality/src/prove_bft/bft.rs:38-565 prove_block_bft; block_finality.rs:200-731; prove_block_data/
signatures.rs:43-141 prove_approvals; header_bphash.rs:34-139 prove_block_header / bp_hash; keys_stakes.rs:18-266;
primitives.rs:32,126,176), reduced to what the prover sees: WHICH proof shapes are produced, in which dependency order.
The reference runs every call one after the other on the CPU; the data flow of its code is this graph:

  stage S  k signature sub-proofs (ed25519-shaped, 2^17 x 234)                      independent  -> sharded over all ranks
  stage C  the signature chain: k recursive_proof steps, strictly sequential (signatures.rs:97-105)   rank 0
  stage A  everything that never reads the signature aggregate -- concurrent with C:
     A1  independent LANES (sequential inside, independent of each other):
           7 block headers (ep2_lb, ep1_fb, b4, b3, b2, b1, bi: bft.rs:69-240), each prove_header_hash
             (header_bphash.rs:34-122): SHA-256 of inner_lite, of inner_rest, hash-of-hashes, with prev_hash, and the
             recursion-bearing proofs around them                                     7 x [4 sha-shaped 2^12 + 5 rec-shaped 2^13]
           next_bp_hash = SHA-256 over the validator list (~110 blocks, header_bphash.rs:125-139)   [3 sha-shaped 2^15]
           prove_eq_array / prove_consecutive_heights (primitives.rs:32,126; bft.rs:124,181,284)    [9 tiny 2^5]
           stake sum over the validator list (keys_stakes.rs:18-266; its inputs, the valid keys, are known natively) [rec-shaped 2^12]
     A2  the 15 recursive aggregations that JOIN lanes but still do not touch the signatures: next_epoch_id (bft.rs:129),
           consecutive heights and Block_i+2 (bft.rs:181-187, block_finality.rs:60-71), epoch_id / bp_hash / Block_n-1 &
           Block_0 (block_finality.rs:337-387), prev_hash / ds / bft hashes and Block_i+1 & Block_i+2 (:463-673)
  stage F  the 5 folds that take the signature aggregate as an input, in the reference's order: signatures + valid keys
           (block_finality.rs:297), + headers (:410), + Block_i+1/i+2 (:691), + the block's own header (:718), public inputs
           (bft.rs:385); then the PoseidonBN128 wrap (src/bin/prove_block.rs:279-300)                  rank 0

With N > 1 ranks stage A runs on ranks 1..N-1 WHILE rank 0 walks the chain (A1 lanes round-robin, one gather among those
ranks, A2 on rank 1); with one rank it runs on a second proving context of the same GPU.  Collectives that move proof bytes:
after S (sub-proofs to rank 0 for the chain), inside A (lane results to the rank that joins them) and after A.  This module
only describes and schedules; the proofs are made by b200_prove through CircuitData.prove.  The aggregation TREE is the
reference's: only independent calls are reordered, no circuit changes.
"""
from . import sharding

N_HEADERS = 7
HEADER_LANE = [("sha", 12)] * 4 + [("rec", 13)] * 5
LEAF_SHAPES = [("sha", 12)] * (4 * N_HEADERS) + [("sha", 15)] * 3 + [("tiny_pi", 5)] * 9 + [("rec", 12)]     # the 41 non-recursive proofs
FINAL_AGGREGATIONS = 55                 # recursive aggregations of a block besides the chain
HEADER_AGGREGATIONS = 5 * N_HEADERS     # inside the header lanes
POST_CHAIN_FOLDS = 5                    # stage F
JOIN_AGGREGATIONS = FINAL_AGGREGATIONS - HEADER_AGGREGATIONS - POST_CHAIN_FOLDS      # stage A2

def lanes():
    """stage A1: [(lane_id, [(task_id, kind, degree_bits), ...])]; task ids are unique over the whole block"""
    out, tid = [], 0
    def lane(shapes):
        nonlocal tid
        tasks = [(tid + i, k, d) for i, (k, d) in enumerate(shapes)]
        tid += len(shapes)
        out.append((len(out), tasks))
    for _ in range(N_HEADERS): lane(HEADER_LANE)
    lane([("sha", 15)] * 3)
    lane([("tiny_pi", 5)] * 9)
    lane([("rec", 12)])
    return out

def num_lane_tasks():
    return sum(len(t) for _, t in lanes())

def join_tasks():
    """stage A2: JOIN_AGGREGATIONS rec-shaped proofs, sequential, on the joining rank"""
    base = num_lane_tasks()
    return [(base + i, "rec", 13) for i in range(JOIN_AGGREGATIONS)]

def stage_a_ranks(world):
    return [0] if world == 1 else list(range(1, world))

def join_rank(world):
    return 0 if world == 1 else 1

def assign_lanes(world, rank):
    """A1 lanes of `rank`: heaviest lanes first, round-robin over the stage-A ranks (rank 0 is busy with the chain when N > 1)"""
    ranks = stage_a_ranks(world)
    if rank not in ranks: return []
    ordered = sorted(lanes(), key=lambda l: -sum(1 << d for _, _, d in l[1]))
    return [l for j, l in enumerate(ordered) if ranks[j % len(ranks)] == rank]

def max_stage_a_tasks_per_rank(world):
    """rows of the gather that closes stage A (lane results + the joins on the joining rank)"""
    m = 0
    for r in range(world):
        n = sum(len(t) for _, t in assign_lanes(world, r)) + (JOIN_AGGREGATIONS if r == join_rank(world) else 0)
        m = max(m, n)
    return m

def max_lane_tasks_per_rank(world):
    return max(sum(len(t) for _, t in assign_lanes(world, r)) for r in range(world))

def summary(k):
    return {"signature_proofs": k, "chain_recursions": k, "leaf_proofs": len(LEAF_SHAPES),
            "aggregations_in_header_lanes": HEADER_AGGREGATIONS, "join_aggregations": JOIN_AGGREGATIONS,
            "post_chain_folds": POST_CHAIN_FOLDS, "wrap_proofs": 1,
            "leaf_shapes": {"sha 2^12": 4 * N_HEADERS, "sha 2^15": 3, "tiny 2^5": 9, "stake (rec-shaped) 2^12": 1}}
