#!/bin/bash
# determinism sweep over kernel buckets: single-GEMM path (k_ty) and two-GEMM path (k_fused); usage: flaky_sweep.sh reps N...
reps=$1; shift
for n in "$@"; do
  a=$(timeout 200 python tools/flaky_bufs.py $n $reps < /dev/null 2>&1 | tail -1)
  b=$(MAMBO_TWO_GEMM=1 timeout 200 python tools/flaky_bufs.py $n $reps < /dev/null 2>&1 | tail -1)
  echo "k_ty: $a | k_fused: $b"
done
