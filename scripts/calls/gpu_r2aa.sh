#!/bin/bash
# Round 2, GPU call AA (1 GPU): programmatic dependent launch along the step's and the plan's kernel chains.
O=gpurun_out
mkdir -p $O
rm -f $O/parity_bench_shapes.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2aa_tests.log 2>&1
echo "tests rc=$?" >> $O/r2aa_tests.log
run() { name=$1; shift; timeout 300 "$@" > $O/r2aa_$name.json 2> $O/r2aa_$name.err; }
for rep in 1 2; do
run ml20m_pdl1_$rep python bench.py --no-cpu --steps 1000
run ml20m_pdl0_$rep python bench.py --no-cpu --steps 1000 --tune pdl=0
done
run side_pdl1 python bench.py --no-cpu --steps 300 --workload sideinfo
run side_pdl0 python bench.py --no-cpu --steps 300 --workload sideinfo --tune pdl=0
run cached_pdl1 python bench.py --no-cpu --steps 1000 --plan cached
run cached_pdl0 python bench.py --no-cpu --steps 1000 --plan cached --tune pdl=0
timeout 300 python scripts/graph_overheads.py ml20m > $O/r2aa_graph_overheads.txt 2>&1
echo done
