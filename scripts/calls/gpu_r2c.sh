#!/bin/bash
# Round 2, GPU call C: tests; two-level cut-row finisher; k_score hoisting A/B; closed-form and fraction bench lines.
O=gpurun_out
mkdir -p $O
rm -f $O/parity_bench_shapes.jsonl
timeout 1500 python -m pytest tests -m gpu -q > $O/r2c_tests.log 2>&1
echo "tests rc=$?" >> $O/r2c_tests.log
timeout 300 python bench.py --no-cpu --steps 1000 > $O/r2c_bench.json 2> $O/r2c_bench.err
VFMB_VARIANT=nohoist timeout 300 python bench.py --no-cpu --steps 1000 > $O/r2c_bench_nohoist.json 2> $O/r2c_bench_nohoist.err
timeout 300 python bench.py --no-cpu --steps 1000 --plan cached > $O/r2c_bench_cached.json 2> $O/r2c_bench_cached.err
timeout 300 python bench.py --no-cpu --steps 1000 --tune fuse_score=1 > $O/r2c_bench_fuse1.json 2> $O/r2c_bench_fuse1.err
timeout 300 python bench.py --steps 200 --workload ml100k > $O/r2c_bench_ml100k.json 2> $O/r2c_bench_ml100k.err
timeout 300 python bench.py --steps 200 --workload fraction > $O/r2c_bench_fraction.json 2> $O/r2c_bench_fraction.err
python scripts/prof_step.py --reserve 1 > $O/r2c_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 78 -c 52 --csv --log-file $O/r2c_launches_warm.csv python scripts/prof_step.py --reserve 1 > $O/r2c_ncu.log 2>&1
echo done
