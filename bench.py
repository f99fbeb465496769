#!/usr/bin/env python3
"""Headline benchmark: ed25519-signature-shaped plonky2 sub-proofs per second on N B200s (BASELINE.json metric).

A "step" = one complete proof (b200_prove: commitments, quotient, openings, FRI, PoW, queries) of the sig-shaped
circuit (wide_ecc_config: 234 wires, 80 routed, 2^17 rows, u32 gate mix) per rank.  With N > 1 the independent
signature sub-proofs are sharded one per rank per step (no data-path collective) and ONE NCCL all-gather of the
per-proof digest records closes the batch, as the reference's NATS fan-in does (signatures.rs:144-274).

  value : proofs/s with the witness already resident in HBM (b200_prove_dev)
  e2e   : proofs/s through the host-buffer C-ABI call (b200_prove): pinned host wires -> H2D, proof bytes -> host
  roofline : the dominant kernel (Poseidon leaf hashing of the 234-column wires LDE), timed with CUDA events on the
             context stream inside the timed region (b200_ctx_enable_timing)
  cpu_baseline / --impl reference : the C oracle prover (a port of upstream plonky2's prove(); the Rust reference
             cannot be built here) on the host cores, on a bounded sample (smaller degree, scaled linearly in rows).
"""
import argparse, json, os, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zk-light-client-implementation_b200"))

METRIC = "ed25519_sig_proofs_per_sec"
UNIT = "proofs/s"

def workload_desc(log_n):
    return {"workload": f"sig-shaped plonky2 proof (wide_ecc_config: 234 wires, 80 routed, degree_bits {log_n}, "
                        "u32 arithmetic/add-many/subtraction/range-check/interleave/comparison + arithmetic/base-sum/"
                        "constant/public-input/poseidon gates; rate_bits 3, cap_height 4, 28 queries, 16-bit PoW)",
            "degree_bits": log_n, "num_wires": 234, "proofs_per_step_per_gpu": 1,
            "l2": "inputs larger than L2: the witness is 8*234*2^degree_bits bytes (245 MB at 2^17) and every pass streams GBs"}

def _host_threads():
    """threads the CPU legs may use: the cores this process is allowed on (not os.cpu_count(): an affinity mask smaller than
    the machine with one OpenMP thread per *machine* core oversubscribes and the spinning waiters make timings 10x unstable)."""
    try: return max(1, len(os.sched_getaffinity(0)))
    except Exception: return os.cpu_count() or 1

def _omp_env(threads):
    env = {"OMP_NUM_THREADS": str(threads), "OMP_WAIT_POLICY": "passive", "OMP_PROC_BIND": "false", "OMP_DYNAMIC": "false"}
    os.environ.update(env)          # read by libgomp when liboracle.so is first loaded (before any oracle import)
    return env

def cpu_prover(kind, log_n, seed, hasher="gl", threads=None):
    """The CPU arm: the C oracle prover (a port of upstream plonky2's prove(); the Rust reference cannot be built here) on the
    circuit workloads.build(kind, log_n, seed) -- the SAME circuit and witness the GPU arm proves.  Returns (prove_fn, info);
    the build-side commitment (OracleCircuit) is outside every timed region, as b200_circuit_upload is in the GPU arm."""
    threads = threads or _host_threads()
    env = _omp_env(threads)
    from b200_plonky2 import workloads          # pure numpy: does not map libb200plonky2.so (b200_plonky2/__init__.py)
    from oracle import prover as oprover
    cd = workloads.build(kind, log_n, seed=seed, hasher=hasher)
    t0 = time.time()
    oc = oprover.OracleCircuit(cd.common, cd.constants_sigmas)
    t_build = time.time() - t0
    def prove():
        t = time.time(); pb = oc.prove(cd.wires, cd.public_inputs); return time.time() - t, pb
    return prove, cd, {"threads": threads, "omp_env": env, "constants_sigmas_commit_s": round(t_build, 3)}

def run_reference(args):
    """--impl reference: the CPU prover on the box's host cores, on the headline circuit itself (not a smaller sample):
    every step is one complete 2^log_n x 234 proof.  One untimed warm-up proof; if K full proofs do not fit the time budget
    the arm times as many as fit and says so in `steps` (never extrapolates)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import hashlib
    prove, cd, info = cpu_prover("sig", args.log_n, seed=1)
    t_warm, pb = prove()                                   # warm-up (page faults of the first large allocations, thread team start)
    steps = args.steps if args.steps * t_warm <= args.ref_budget_s else max(1, int(args.ref_budget_s // t_warm))
    times = []
    for _ in range(steps):
        dt, pb = prove(); times.append(dt)
    total = sum(times)
    v = steps / total
    # the other proof shapes of a block, for a CPU figure beside bft_finality_block (composed from measured per-proof times)
    block = None
    if not args.no_block:
        try:
            p_rec, _, _ = cpu_prover("rec", 13, seed=3); p_rec(); t_rec = min(p_rec()[0] for _ in range(2))
            p_wrap, _, _ = cpu_prover("rec", 12, seed=3, hasher="bn254"); t_wrap = p_wrap()[0]
            k = args.block_signatures
            block = {"bft_block_ms": round((k * total / steps + (k + 55) * t_rec + t_wrap) * 1e3, 1), "composed": True,
                     "sig_proof_ms": round(total / steps * 1e3, 1), "recursive_proof_2^13_ms": round(t_rec * 1e3, 1), "bn128_wrap_2^12_ms": round(t_wrap * 1e3, 1),
                     "signatures": k, "recursive_proofs": k + 55,
                     "note": "k sig-shaped + (k+55) recursive_proof-shaped + 1 BN128-wrap proofs, each timed on this host (all cores per proof, one at a "
                             "time) and summed -- the whole block is not run end to end on the CPU (it would take > 10 min)"}
        except Exception as e:
            block = {"bft_block_ms": None, "error": str(e)}
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": 1,
           "ms_per_step": 1000.0 * total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
           "data": "synthetic", "config": workload_desc(args.log_n),
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": info["threads"], "kind": "port",
                            "sample": f"{steps} complete oracle-prover proofs of the headline circuit itself (degree_bits {args.log_n}, 234 wires, seed 1), "
                                      f"{total / steps:.2f} s each (min {min(times):.2f}, max {max(times):.2f}); warm-up proof {t_warm:.2f} s; "
                                      "a C/OpenMP port of upstream plonky2's prove(), NOT the Rust prover (no toolchain in the image)",
                            "step_seconds": [round(t, 3) for t in times], **info},
           "proof_sha256": hashlib.sha256(pb).hexdigest(), "proof_bytes": len(pb),
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if steps != args.steps:
        out["steps_note"] = f"{args.steps} steps requested; {steps} complete proofs fit the {args.ref_budget_s:.0f} s budget of this arm"
    if block is not None: out["bft_finality_block"] = block
    print(json.dumps(out))

class ClockSampler(threading.Thread):
    def __init__(self, dev):
        super().__init__(daemon=True); self.dev = dev; self.rows = []; self.stop = False
    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop:
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.dev)],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o: self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)
    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}

def run_b200(args):
    import numpy as np, torch
    import b200_plonky2 as bp
    from b200_plonky2 import workloads, sharding, lib, c_void_p, c_size_t
    import ctypes
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    real_stdout = os.dup(1)
    os.dup2(2, 1)                     # NCCL prints its version banner on stdout: keep stdout for the ONE JSON line
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        t = torch.zeros(8, dtype=torch.int64, device="cuda")
        dist.all_gather([torch.empty_like(t) for _ in range(world)], t)     # communicator set-up outside the timed region
        torch.cuda.synchronize()
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    ctx = bp.Context(local); ctx.set_stream(stream.cuda_stream)

    cd = workloads.build("sig", args.log_n, seed=1)
    t_b0 = time.perf_counter(); cache_hit0 = cd.build_on_device(ctx); build_cold_ms = (time.perf_counter() - t_b0) * 1e3     # b200_circuit_build: sigma / selector polynomials + commitment on the GPU
    W, n = cd.wires.shape
    host_wires = torch.from_numpy(cd.wires).pin_memory()
    dev_wires = host_wires.to("cuda", non_blocking=True)
    pis = np.array(cd.public_inputs, dtype=np.uint64)
    out = ctypes.POINTER(ctypes.c_uint8)(); ln = c_size_t()
    proof_len = [0]
    def prove(dev, full=False):
        fn = lib.b200_prove_dev if dev else lib.b200_prove
        ptr = dev_wires.data_ptr() if dev else host_wires.data_ptr()
        rc = fn(ctx.h, cd._handle, c_void_p(ptr), pis.ctypes.data_as(c_void_p), len(pis), ctypes.c_uint64(0), ctypes.byref(out), ctypes.byref(ln))
        if rc: raise RuntimeError(lib.b200_last_error(ctx.h).decode())
        b = ctypes.string_at(out, ln.value if full else 64); proof_len[0] = ln.value
        lib.b200_free(out)
        return b
    def barrier():
        torch.cuda.synchronize()
        if dist: dist.barrier()
        torch.cuda.synchronize()
    # extra proving contexts on this GPU (one host thread + stream each; the C ABI's threading model): independent proofs
    # overlap one proof's latency-bound tails (upper Merkle levels, FRI layers, Fiat-Shamir round trips) with another's
    # throughput-bound kernels.  Used by the whole-block run and reported as `concurrent_streams`; the headline timed
    # region above stays one proof at a time on one stream.
    import threading, copy
    pool = []          # [(ctx, circuit)] beyond the main context
    def pool_ensure(S):
        while len(pool) < S - 1:
            c = bp.Context(local); d = copy.copy(cd); d._handle = None; d.upload(c)
            pool.append((c, d))
            prove_on(len(pool), False)
    def prove_on(slot, dev, full=False):
        if slot == 0: return prove(dev, full)
        c, d = pool[slot - 1]
        o = ctypes.POINTER(ctypes.c_uint8)(); l2 = c_size_t()
        fn = lib.b200_prove_dev if dev else lib.b200_prove
        ptr = dev_wires.data_ptr() if dev else host_wires.data_ptr()
        rc = fn(c.h, d._handle, c_void_p(ptr), pis.ctypes.data_as(c_void_p), len(pis), ctypes.c_uint64(0), ctypes.byref(o), ctypes.byref(l2))
        if rc: raise RuntimeError(lib.b200_last_error(c.h).decode())
        b = ctypes.string_at(o, l2.value if full else 64); lib.b200_free(o)
        return b
    def prove_many(indices, dev, S, full=False):
        """prove one sig-shaped proof per entry of `indices` on S concurrent contexts; returns {index: proof bytes (head only unless full)}"""
        if S <= 1 or len(indices) <= 1: return {i: prove(dev, full) for i in indices}
        pool_ensure(S)
        heads, errs = {}, []
        def work(slot):
            try:
                for i in indices[slot::S]: heads[i] = prove_on(slot, dev, full)
            except Exception as e: errs.append(e)
        ths = [threading.Thread(target=work, args=(sl,)) for sl in range(S)]
        for t in ths: t.start()
        for t in ths: t.join()
        if errs: raise errs[0]
        return heads
    def timed(dev, steps, gather):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        mine = {}
        for s in range(steps):
            mine[s * world + rank] = prove(dev, full=True)                     # signature index s*world+rank -> this rank
        if gather:
            # the single NVLink collective: the COMPLETE sub-proofs (what recursive_proof consumes: signatures.rs:97-105,225-232)
            allp = sharding.gather_proofs(mine, dist, "cuda", steps, proof_len[0])
            assert sorted(allp) == list(range(steps * world)) and all(len(b) == proof_len[0] for b in allp.values())
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if dist:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        return ms

    for _ in range(args.warmup): prove(True)
    # warm torch's own CUDA kernels and the collective AT THE SHAPE OF THE TIMED REGION (NCCL sets up its buffers per message
    # size on first use: ~40 ms once, which a first timed region would otherwise carry)
    _w = prove(True, full=True)
    sharding.gather_proofs({s_ * world + rank: _w for s_ in range(args.steps)}, dist, "cuda", args.steps, proof_len[0])
    torch.cuda.synchronize()
    sampler = ClockSampler(local); sampler.start()
    ctx.enable_timing(True); ctx.timings()
    l0 = ctx.kernel_launches()
    ms_dev = timed(True, args.steps, True)
    launches = ctx.kernel_launches() - l0
    phases = ctx.timings()
    ctx.enable_timing(False)
    for _ in range(min(args.warmup, 1)): prove(False)
    ms_e2e = timed(False, args.steps, True)
    sampler.stop = True; sampler.join(timeout=2)

    # ---- the proofs the timed loops produce are checked at the headline size, on every rank: the product's own verifier
    # (b200_verify, host only) must accept the device-resident and the host-buffer proof, and both must be the same bytes
    import hashlib
    proof_dev, proof_host = prove(True, full=True), prove(False, full=True)
    ok_verify = int(bool(cd.verify(proof_dev)) and proof_dev == proof_host)
    if dist:
        fl = torch.tensor([ok_verify], dtype=torch.int64, device="cuda"); dist.all_reduce(fl, op=dist.ReduceOp.MIN); ok_verify = int(fl.item())
    if not ok_verify:
        raise SystemExit("bench: a proof of the headline circuit was rejected by b200_verify (or dev/host proofs differ) -- no number reported")

    # ---- one whole BFT-finality block, measured (second half of BASELINE.json's metric; configs[1] at N=1, configs[2] at N=8).
    # Task graph: b200_plonky2/block_dag.py (the data flow of prove_bft/bft.rs:38-565 + block_finality.rs:200-731).
    #   S: k sig-shaped sub-proofs sharded over the ranks (host witnesses through b200_prove) -> all-gather of the complete proofs
    #   C: the k-step recursive chain on rank 0, WHILE stage A runs on ranks 1..N-1 (N = 1: on a second context of the same GPU):
    #      A1 = independent lanes (7 headers x [4 SHA-256-shaped + 5 recursive], bp_hash, tiny proofs, stake), a gather of the lane
    #      results among those ranks, A2 = the 15 recursive aggregations that join lanes without touching the signatures
    #   -> all-gather of stage A's proofs -> F: the 5 folds that consume the signature aggregate + the PoseidonBN128 wrap on rank 0.
    # Wall clock between barriers, max over ranks.
    block = None
    if not args.no_block:
        try:
            from b200_plonky2 import block_dag
            k = args.block_signatures
            rec_cd = wrap_cd = None
            ready = 1
            leaf_ctx = ctx if world > 1 else bp.Context(local)          # N = 1: stage A gets its own context / stream / host thread
            leaf_ctx2 = None if world > 1 else bp.Context(local)        # ... and a second one: lanes are independent, small proofs are latency-bound
            my_lanes = block_dag.assign_lanes(world, rank)
            a_ranks, j_rank = block_dag.stage_a_ranks(world), block_dag.join_rank(world)
            my_joins = block_dag.join_tasks() if rank == j_rank else []
            sub = dist.new_group(ranks=a_ranks) if (dist and len(a_ranks) > 1) else None      # collective call: every rank makes it
            leaf_cds, leaf_cds2 = {}, {}
            try:
                for kind, db in sorted({(kd, d) for _, tasks in my_lanes for _, kd, d in tasks} | {(kd, d) for _, kd, d in my_joins}):
                    c_ = workloads.build(kind, db, seed=5); c_.upload(leaf_ctx); c_.prove(leaf_ctx); leaf_cds[(kind, db)] = c_
                    if leaf_ctx2 is not None:
                        c2 = workloads.build(kind, db, seed=5); c2.upload(leaf_ctx2); c2.prove(leaf_ctx2); leaf_cds2[(kind, db)] = c2
                if rank == 0:
                    rec_cd = workloads.build("rec", 13, seed=3); rec_cd.upload(ctx)
                    wrap_cd = workloads.build("rec", 12, seed=3, hasher="bn254"); wrap_cd.upload(ctx)
                    rec_w = ctx.pinned_copy(rec_cd.wires); wrap_w = ctx.pinned_copy(wrap_cd.wires)
                    rec_cd.prove(ctx, wires=rec_w); wrap_cd.prove(ctx, wires=wrap_w)
            except Exception as e:
                ready = 0; print(f"[bench] block set-up failed on rank {rank}: {e}", file=sys.stderr)
            if dist:      # every rank takes the same branch: a one-sided failure must not leave the others in a barrier
                fl = torch.tensor([ready], dtype=torch.int64, device="cuda"); dist.all_reduce(fl, op=dist.ReduceOp.MIN); ready = int(fl.item())
            if not ready: raise RuntimeError("block set-up failed")
            mine = sharding.shard_indices(k, world, rank)
            if args.block_streams > 1: pool_ensure(args.block_streams)
            leaf_stride = 200000
            # warm the proof collectives at their real shapes (communicator buffers) outside the timed region
            sharding.gather_proofs({}, dist, "cuda", (k + world - 1) // world, proof_len[0])
            if sub is not None and rank in a_ranks:
                sharding.gather_proofs({}, dist, "cuda", block_dag.max_lane_tasks_per_rank(world), leaf_stride, group=sub, group_size=len(a_ranks))
            barrier(); t0 = time.perf_counter()
            # S
            sig = prove_many(mine, False, args.block_streams, full=True)
            t_sig = time.perf_counter() - t0
            allsig = sharding.gather_proofs(sig, dist, "cuda", (k + world - 1) // world, proof_len[0])
            assert sorted(allsig) == list(range(k))
            t_g1 = time.perf_counter() - t0
            # C (rank 0) || A (ranks 1..N-1, or a second context)
            a_out, errs, t_a_busy, t_a_gather = {}, [], [0.0], [0.0]
            def run_stage_a():
                try:
                    ta = time.perf_counter()
                    def run_lanes(lanes_, cds_, cx_):
                        try:
                            for _, tasks in lanes_:                           # a lane is sequential; lanes are independent
                                for tid_, kind, db in tasks: a_out[tid_] = cds_[(kind, db)].prove(cx_)
                        except Exception as e: errs.append(e)
                    if leaf_ctx2 is not None:
                        t2 = threading.Thread(target=run_lanes, args=(my_lanes[1::2], leaf_cds2, leaf_ctx2)); t2.start()
                        run_lanes(my_lanes[0::2], leaf_cds, leaf_ctx); t2.join()
                    else: run_lanes(my_lanes, leaf_cds, leaf_ctx)
                    if sub is not None:                                       # lane results -> the joining rank (and its peers)
                        tg = time.perf_counter()
                        lanes_all = sharding.gather_proofs(dict(a_out), dist, f"cuda:{local}", block_dag.max_lane_tasks_per_rank(world), leaf_stride,
                                                           group=sub, group_size=len(a_ranks))   # explicit device: torch's current device is per thread
                        assert len(lanes_all) == block_dag.num_lane_tasks()
                        t_a_gather[0] = time.perf_counter() - tg
                    for tid_, kind, db in my_joins: a_out[tid_] = leaf_cds[(kind, db)].prove(leaf_ctx)      # A2, in dependency order
                    t_a_busy[0] = time.perf_counter() - ta
                except Exception as e: errs.append(e)
            th = None
            if rank in a_ranks: th = threading.Thread(target=run_stage_a); th.start()
            t_chain = 0.0
            if rank == 0:
                tc = time.perf_counter()
                for i in range(k): rec_cd.prove(ctx, wires=rec_w)            # chain step i consumes sub-proof i (validator order)
                t_chain = time.perf_counter() - tc
            if th is not None: th.join()
            if errs: raise errs[0]
            t_ca = time.perf_counter() - t0
            all_a = sharding.gather_proofs(a_out, dist, "cuda", block_dag.max_stage_a_tasks_per_rank(world), leaf_stride)
            assert sorted(all_a) == list(range(block_dag.num_lane_tasks() + block_dag.JOIN_AGGREGATIONS))
            t_g2 = time.perf_counter() - t0
            # F
            if rank == 0:
                for _ in range(block_dag.POST_CHAIN_FOLDS): rec_cd.prove(ctx, wires=rec_w)
                wrap_cd.prove(ctx, wires=wrap_w)
            barrier(); t_all = time.perf_counter() - t0
            tt = torch.tensor([t_sig, t_g1, t_ca, t_g2, t_all, t_chain, t_a_busy[0], t_a_gather[0]], dtype=torch.float64, device="cuda")
            if dist: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_sig, t_g1, t_ca, t_g2, t_all, t_chain, t_leaf, t_ag = (float(x) for x in tt.tolist())
            block = {"bft_block_ms": round(t_all * 1e3, 1), "signature_stage_ms": round(t_sig * 1e3, 1),
                     "subproof_all_gather_ms": round((t_g1 - t_sig) * 1e3, 2),
                     "subproof_all_gather_bytes": sharding.payload_bytes((k + world - 1) // world, proof_len[0], world),
                     "chain_and_stage_a_ms": round((t_ca - t_g1) * 1e3, 1), "chain_ms_rank0": round(t_chain * 1e3, 1),
                     "stage_a_busy_ms_max_rank": round(t_leaf * 1e3, 1), "stage_a_lane_gather_ms": round(t_ag * 1e3, 2),
                     "stage_a_all_gather_ms": round((t_g2 - t_ca) * 1e3, 2),
                     "post_chain_folds_and_wrap_ms": round((t_all - t_g2) * 1e3, 1),
                     "signatures": k, "sig_proofs_per_rank_max": len(sharding.shard_indices(k, world, 0)),
                     "streams_per_gpu_in_signature_stage": args.block_streams, "task_graph": block_dag.summary(k),
                     "stage_a_placement": "ranks 1..N-1 while rank 0 walks the chain (joins on rank 1)" if world > 1 else "two more contexts on the same GPU (lanes split between them), concurrent with the chain",
                     "note": "measured end to end between barriers (max over ranks): prover time of every proof of the block's task graph with host witnesses; "
                             "witness generation and circuit builds stay in Rust and are not included; the recursive steps re-prove one recursion-shaped witness; "
                             "only calls whose inputs never include the signature aggregate run beside the chain, the aggregation tree is the reference's"}
            if rank == 0: rec_cd.free(); wrap_cd.free()
            for c_ in list(leaf_cds.values()) + list(leaf_cds2.values()): c_.free()
            if leaf_ctx is not ctx: leaf_ctx.close()
            if leaf_ctx2 is not None: leaf_ctx2.close()
        except Exception as e:
            block = {"bft_block_ms": None, "error": str(e)}

    conc = None
    if args.concurrent_streams > 1:
        try:
            S = args.concurrent_streams
            pool_ensure(S)
            res = {}
            for name, dev in (("value", True), ("e2e", False)):
                prove_many(list(range(S)), dev, S)
                barrier(); t0 = time.perf_counter()
                prove_many(list(range(S * args.steps)), dev, S)
                barrier(); dt = time.perf_counter() - t0
                tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
                if dist: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                res[name] = S * args.steps * world / float(tt.item())
            conc = {"streams_per_gpu": S, "value": res["value"], "e2e": res["e2e"], "unit": UNIT,
                    "note": "same proofs issued from S host threads per GPU, each with its own context/stream (wall clock between barriers, "
                            "max over ranks); reported beside the one-proof-at-a-time headline, not instead of it"}
        except Exception as e:
            conc = {"error": str(e)}

    total_proofs = args.steps * world
    value = total_proofs / (ms_dev * 1e-3)
    e2e_value = total_proofs / (ms_e2e * 1e-3)
    if rank != 0:
        if dist: dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel: Poseidon leaf hashing of the wires LDE ----
    peaks = {}
    try: peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception: pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    L = n << 3
    leaf = phases.get("leaf_hash_wide", [0.0, 0])
    # per prove: the wide leaf-hash phase covers the wires oracle (and nothing else has >= 64 columns)
    per_launch_ms = leaf[0] / max(1, leaf[1])
    alg_bytes = 8.0 * L * W + 32.0 * L
    achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms else None
    perms = L * ((W + 7) // 8)
    # instruction-issue view of the same kernel (what actually bounds it): thread-instructions per permutation from the committed ncu
    # capture of this kernel and shape, times the permutations per launch, over the LIVE launch time, against the SM issue rate
    # (4 warp-instructions per clock per SM at the clock seen during the run)
    NCU_INSTR_PER_PERM = 23696          # profiles/ncu_k_leaves_colmajor_r02_v1.json: 23.30 G warp-instructions / 31 457 280 permutations
    clk = sampler.summary().get("sm_mhz") or 1965
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    issue_peak = sms * 128.0 * clk * 1e6                     # thread-instructions per second
    issue_ach = perms * NCU_INSTR_PER_PERM / (per_launch_ms * 1e-3) if per_launch_ms else None
    lde = phases.get("lde", [0.0, 0]); intt = phases.get("intt", [0.0, 0])
    lde_cols = W + 2 * (1 + cd.common["num_partial_products"]) + 2 * cd.common["quotient_degree_factor"]
    lde_bytes = 72.0 * n * lde_cols                          # 8n read + 64n written per column, every LDE of one proof
    roofline = {"bound": "hbm", "kernel": "k_leaves_colmajor (Poseidon sponge over the 234-column wires LDE, 2^20 leaves)",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": (achieved / hbm_peak) if achieved else None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "traffic": 2.0040e9 if args.log_n == 17 else None,
                "traffic_source": "profiles/ncu_k_leaves_colmajor_r02_v1.json: dram read 1.967 GB + write 0.0369 GB per launch (ncu --set full, degree_bits 17)",
                "ms_per_launch": per_launch_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "share_of_step": (leaf[0] / ms_dev) if ms_dev else None,
                "instr_per_perm": NCU_INSTR_PER_PERM,
                "issue_frac": (issue_ach / issue_peak) if issue_ach else None,
                "issue": {"achieved_Tinstr_s": issue_ach / 1e12 if issue_ach else None, "peak_Tinstr_s": issue_peak / 1e12,
                          "peak_source": f"{sms} SMs x 4 warp-instructions/clk x 32 lanes x {clk} MHz (clock sampled during the timed region)",
                          "instr_source": "profiles/ncu_k_leaves_colmajor_r02_v1.json (smsp__inst_executed.sum, same kernel and shape); round 1: 33 200"},
                "alu_note": {"permutations_per_launch": perms, "Mperm_per_s": perms / (per_launch_ms * 1e-3) / 1e6 if per_launch_ms else None,
                             "ncu": {"source": "profiles/ncu_k_leaves_colmajor_r02_v1.json (ncu --set full, same kernel and shape)",
                                     "issue_slots_active_pct": 74.0, "pipe_fmaheavy_active_pct": 82.2, "pipe_alu_active_pct": 68.8,
                                     "instructions_per_permutation": NCU_INSTR_PER_PERM, "dram_throughput_pct_of_peak": 0.89,
                                     "top_stalls": {"math_pipe_throttle": 0.33, "not_selected": 0.32}},
                             "comment": "bound by integer instruction issue / the FMA-heavy pipe (IMAD.WIDE.U32), not by HBM; see DESIGN.md section 3"},
                "other_kernels": {
                    "ntt_lde": {"kernel": "k_ntt2_pass<1> + k_ntt2_pass<2> (x8 coset LDE of all four polynomial batches of a proof, bulk-copy staged)",
                                "achieved": lde_bytes / (lde[0] / args.steps * 1e-3) / 1e9 if lde[0] else None, "peak": hbm_peak, "unit": "GB/s",
                                "frac": lde_bytes / (lde[0] / args.steps * 1e-3) / 1e9 / hbm_peak if lde[0] else None,
                                "algorithmic_bytes_per_step": lde_bytes, "ms_per_step": lde[0] / args.steps if lde[0] else None, "columns": lde_cols},
                    "ntt_intt": {"kernel": "k_ntt2_pass<0> + k_ntt2_pass<2> (inverse NTT of the wire and Z / partial-product columns)",
                                 "achieved": 16.0 * n * (W + 2 * (1 + cd.common["num_partial_products"])) / (intt[0] / args.steps * 1e-3) / 1e9 if intt[0] else None,
                                 "peak": hbm_peak, "unit": "GB/s", "ms_per_step": intt[0] / args.steps if intt[0] else None}}}
    out_json = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
                "data": "synthetic", "config": workload_desc(args.log_n),
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": int(W * n * 8 + len(pis) * 8), "d2h_bytes_per_step": int(proof_len[0])},
                "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roofline,
                "phases_ms_per_step": {k: round(v[0] / args.steps, 4) for k, v in phases.items()},
                "proof_bytes": proof_len[0]}
    if block is not None: out_json["bft_finality_block"] = block
    if conc is not None: out_json["concurrent_streams"] = conc
    if world == 1:
        # CircuitBuilder::build()'s device half (b200_circuit_build): cold = sigma / selector polynomial generation + iNTT + LDE + Merkle of
        # the 86 constant / sigma columns + digest; cached = hashing the build inputs and finding the resident circuit (what every
        # repeated recursive_proof / per-signature build of the reference becomes: recursion.rs:36-96, ed25519.rs:66-77)
        try:
            import copy as _copy
            again = _copy.copy(cd); again._handle = None
            t_ = time.perf_counter(); hit = again.build_on_device(ctx); cached_ms = (time.perf_counter() - t_) * 1e3
            rec_b = workloads.build("rec", 13, seed=3)
            t_ = time.perf_counter(); rec_b.build_on_device(ctx); rec_cold = (time.perf_counter() - t_) * 1e3
            rec_b2 = workloads.build("rec", 13, seed=3)
            t_ = time.perf_counter(); hit2 = rec_b2.build_on_device(ctx); rec_cached = (time.perf_counter() - t_) * 1e3
            out_json["circuit_build"] = {"sig_2^%d_cold_ms" % args.log_n: round(build_cold_ms, 2), "sig_cached_ms": round(cached_ms, 2), "sig_cache_hit": bool(hit),
                                         "recursive_2^13_cold_ms": round(rec_cold, 2), "recursive_2^13_cached_ms": round(rec_cached, 2), "recursive_cache_hit": bool(hit2),
                                         "cache": ctx.cache_stats(),
                                         "note": "host wall clock of b200_circuit_build incl. H2D of the gate placement / copy-class successor map; outside the timed proving loops"}
            again.free(); rec_b.free(); rec_b2.free()
        except Exception as e:
            out_json["circuit_build"] = {"error": str(e)}
    if world == 1 and not args.no_aux:
        # the other proof shapes of one BFT-finality block (second half of BASELINE.json's metric): a recursive_proof-shaped
        # circuit (13 standard gates, 135 wires, 2^13 rows) and the final PoseidonBN128 wrap (2^12 rows), ms per proof
        aux = {}
        for name, kind, db, hs in (("recursive_proof_2^13_ms", "rec", 13, "gl"), ("bn128_wrap_2^12_ms", "rec", 12, "bn254")):
            try:
                c2 = workloads.build(kind, db, seed=3, hasher=hs); c2.upload(ctx)
                c2.prove(ctx); torch.cuda.synchronize()
                t0 = time.time()
                for _ in range(3): c2.prove(ctx)
                torch.cuda.synchronize()
                aux[name] = round((time.time() - t0) / 3 * 1e3, 3)
                c2.free()
            except Exception as e:
                aux[name] = f"failed: {e}"
        # small proofs are latency-bound (sequential Merkle levels, Fiat-Shamir round trips): independent ones overlap when
        # issued from several host threads, each with its own context/stream (the ABI's threading model)
        try:
            import threading
            nthr, per = 4, 6
            ctxs = [bp.Context(local) for _ in range(nthr)]
            cds = [workloads.build("rec", 13, seed=3) for _ in range(nthr)]
            for c_, d_ in zip(ctxs, cds): d_.upload(c_); d_.prove(c_)
            def work(c_, d_):
                for _ in range(per): d_.prove(c_)
            ths = [threading.Thread(target=work, args=(c_, d_)) for c_, d_ in zip(ctxs, cds)]
            torch.cuda.synchronize(); t0 = time.time()
            for t in ths: t.start()
            for t in ths: t.join()
            torch.cuda.synchronize()
            aux["recursive_proof_2^13_ms_4_concurrent_contexts"] = round((time.time() - t0) / (nthr * per) * 1e3, 3)
            for d_ in cds: d_.free()
            for c_ in ctxs: c_.close()
        except Exception as e:
            aux["recursive_proof_2^13_ms_4_concurrent_contexts"] = f"failed: {e}"
        k = 73
        if all(isinstance(v, float) for v in aux.values()):
            aux["bft_block_estimate_ms"] = round(k * ms_e2e / args.steps + (k + 55) * aux["recursive_proof_2^13_ms"] + aux["bn128_wrap_2^12_ms"], 1)
            aux["bft_block_estimate_note"] = ("k=73 signed approvals (prove_random fixture): k sig-shaped proofs + (k+55) recursive_proof-shaped proofs + 1 BN128 wrap run one after the other, "
                                              "prover time only (no witness generation / circuit builds, which stay in Rust); SHA-256 sub-proofs not included")
        out_json["aux_proof_shapes"] = aux
    out_json["headline_proof_check"] = {"b200_verify": "accepted on every rank", "dev_bytes_equal_host_bytes": True,
                                        "proof_sha256": hashlib.sha256(proof_dev).hexdigest(), "degree_bits": args.log_n}
    if world == 1 and not args.no_cpu_baseline:
        # CPU baseline leg = ONE complete oracle-prover proof of the headline circuit itself (same circuit, same witness), which is
        # also the parity check at the benchmark config: the GPU proof must be byte-identical to it.
        threads = _host_threads(); env = _omp_env(threads)
        try:
            from oracle import prover as oprover
            oc = oprover.OracleCircuit(cd.common, cd.constants_sigmas)
            t0 = time.time(); pb = oc.prove(cd.wires, cd.public_inputs); dt = time.time() - t0
            same = pb == proof_dev
            out_json["cpu_baseline"] = {"value": 1.0 / dt, "unit": UNIT, "cores": threads, "kind": "port", "omp_env": env,
                                        "sample": f"1 complete oracle-prover proof of the headline circuit itself (degree_bits {args.log_n}, 234 wires): {dt:.2f} s "
                                                  "(first proof in this process, no warm-up; the --impl reference arm times warmed-up proofs)",
                                        "sample_seconds": round(dt, 3),
                                        "parity_check": "gpu proof bytes == oracle proof bytes at the headline config" if same else "MISMATCH"}
            out_json["headline_proof_check"]["bytes_equal_oracle_prover"] = bool(same)
            if not same:
                raise SystemExit("bench: GPU proof bytes differ from the oracle prover's at the headline config -- no number reported")
        except SystemExit:
            raise
        except Exception as e:       # the baseline is reported, never the product path
            out_json["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": threads, "kind": "port", "sample": f"failed: {e}"}
    sys.stdout.flush(); os.dup2(real_stdout, 1)
    print(json.dumps(out_json), flush=True)
    if dist: dist.destroy_process_group()

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-n", type=int, default=17)
    ap.add_argument("--ref-budget-s", type=float, default=240.0, help="--impl reference: wall-clock budget of the timed proofs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true")
    ap.add_argument("--no-block", action="store_true", help="skip the measured whole-block run (k sig proofs + fold + wrap)")
    ap.add_argument("--block-streams", type=int, default=2, help="concurrent proving contexts per GPU in the block's signature stage")
    ap.add_argument("--concurrent-streams", type=int, default=3, help="extra measurement: S concurrent contexts per GPU (0/1 = skip)")
    ap.add_argument("--block-signatures", type=int, default=73, help="signed approvals in the block (prove_random fixture: 73)")
    args = ap.parse_args()
    if args.impl == "reference": run_reference(args)
    else: run_b200(args)

if __name__ == "__main__":
    main()
