#!/bin/bash
# Round 2, GPU call AD (1 GPU): final single-GPU state -- full suite, bench lines, warm launch list.
O=gpurun_out
mkdir -p $O
rm -f $O/parity_bench_shapes.jsonl
timeout 1500 python -m pytest tests -m gpu -q > $O/r2ad_tests.log 2>&1
echo "tests rc=$?" >> $O/r2ad_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2ad_bench_driver_args.json 2> $O/r2ad_bench_driver_args.err
timeout 600 python bench.py --no-cpu > $O/r2ad_bench_default.json 2> $O/r2ad_bench_default.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2ad_bench_reference.json 2> $O/r2ad_bench_reference.err
for w in sideinfo ml100k fraction; do
  timeout 300 python bench.py --no-cpu --steps 300 --workload $w > $O/r2ad_bench_$w.json 2> $O/r2ad_bench_$w.err
done
python scripts/prof_step.py --reserve 1 > $O/r2ad_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 78 -c 52 --csv --log-file $O/r2ad_launches_warm.csv python scripts/prof_step.py --reserve 1 > $O/r2ad_ncu.log 2>&1
echo done
