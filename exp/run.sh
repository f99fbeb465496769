#!/bin/bash
# scratch experiment: leaf-hash kernel variants (S-box interleave, resident blocks per SM); not part of the product
cd "$(dirname "$0")/.."
cp zk-light-client-implementation_b200/libb200plonky2.so /tmp/lib_orig.so
for v in u2 u3 u4 b7 b8 u2; do
  cp exp/lib_$v.so zk-light-client-implementation_b200/libb200plonky2.so
  echo -n "$v " ; python tools/microbench.py 17 234 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['poseidon_merkle_ms'], d['lde_commit_ms'])"
done
cp /tmp/lib_orig.so zk-light-client-implementation_b200/libb200plonky2.so
