#!/bin/bash
# Round 2, GPU call O (1 GPU): pairwise own-term moved into the row kernel (F > 2); suite + sideinfo.
O=gpurun_out
mkdir -p $O
rm -f $O/parity_bench_shapes.jsonl
timeout 1500 python -m pytest tests -m gpu -q > $O/r2o_tests.log 2>&1
echo "tests rc=$?" >> $O/r2o_tests.log
timeout 300 python bench.py --no-cpu --steps 300 --workload sideinfo > $O/r2o_bench_sideinfo.json 2> $O/r2o_bench_sideinfo.err
timeout 300 python bench.py --no-cpu --steps 1000 > $O/r2o_bench.json 2> $O/r2o_bench.err
python scripts/prof_step.py --reserve 1 --workload sideinfo --rows 1000000 > $O/r2o_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 78 -c 52 --csv --log-file $O/r2o_launches_sideinfo_warm.csv python scripts/prof_step.py --reserve 1 --workload sideinfo --rows 1000000 > $O/r2o_ncu.log 2>&1
echo done
