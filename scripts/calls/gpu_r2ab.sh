#!/bin/bash
# Round 2, GPU call AB (1 GPU): L2 residency hints for the rows the update re-reads; fused vs unfused eager.
O=gpurun_out
mkdir -p $O
run() { name=$1; shift; timeout 300 "$@" > $O/r2ab_$name.json 2> $O/r2ab_$name.err; }
run keep0 python bench.py --no-cpu --steps 1000
run keep1 python bench.py --no-cpu --steps 1000 --tune l2_keep=1
run keep3 python bench.py --no-cpu --steps 1000 --tune l2_keep=3
run keep7 python bench.py --no-cpu --steps 1000 --tune l2_keep=7
run side_keep0 python bench.py --no-cpu --steps 300 --workload sideinfo
run side_keep1 python bench.py --no-cpu --steps 300 --workload sideinfo --tune l2_keep=1
run cached_fuse0 python bench.py --no-cpu --steps 1000 --plan cached
run cached_fuse1 python bench.py --no-cpu --steps 1000 --plan cached --tune fuse_score=1
timeout 300 python -m pytest tests/test_gpu_sampled.py -m gpu -q -x > $O/r2ab_tests.log 2>&1
echo done
