#!/bin/bash
# Round 2, GPU call AC (1 GPU): k_score prefetches the rows of the update into L2.
O=gpurun_out
mkdir -p $O
run() { name=$1; shift; timeout 300 "$@" > $O/r2ac_$name.json 2> $O/r2ac_$name.err; }
run base python bench.py --no-cpu --steps 1000
run keep1 python bench.py --no-cpu --steps 1000 --tune l2_keep=1
run sp6 python bench.py --no-cpu --steps 1000 --tune score_prefetch=6
run sp7 python bench.py --no-cpu --steps 1000 --tune score_prefetch=7
run sp2 python bench.py --no-cpu --steps 1000 --tune score_prefetch=2
run keep1_sp6 python bench.py --no-cpu --steps 1000 --tune l2_keep=1 --tune score_prefetch=6
run keep1_sp2 python bench.py --no-cpu --steps 1000 --tune l2_keep=1 --tune score_prefetch=2
run side_base python bench.py --no-cpu --steps 300 --workload sideinfo
run side_sp7 python bench.py --no-cpu --steps 300 --workload sideinfo --tune score_prefetch=7
run side_keep1_sp6 python bench.py --no-cpu --steps 300 --workload sideinfo --tune l2_keep=1 --tune score_prefetch=6
echo done
