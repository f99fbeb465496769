"""CPU oracle for the plonky2 prover hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product (zk-light-client-implementation_b200/, include/) imports, links or executes
anything in this package.  Allowed importers: tests/, __graft_entry__.smoke(), and bench.py's
cpu_baseline / --impl reference legs.

Parity status: the *verifier-side* arithmetic (Goldilocks, F_p^2, Poseidon-GL, Poseidon-BN254,
challenger, gate evaluators, FRI verification, wire formats) is PINNED by the reference's own
fixtures and KATs (tests/test_oracle_*.py).  The *prover-side* decisions a verifier cannot see
(PoW nonce choice, ed25519 witness contents, circuit build) are parity-unpinned in this
container: the reference prover is Rust (plonky2 @ wormhole-foundation/plonky2-near 2244a9d),
is not vendored in /root/reference and there is no Rust toolchain here.  See DESIGN.md.
"""
