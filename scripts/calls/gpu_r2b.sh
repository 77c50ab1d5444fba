#!/bin/bash
# Round 2, GPU call B: full GPU test suite (cut rows finished inside the gather kernels), bandwidth
# ceiling tool, bench, warm-cache launch list of the eager step.
O=gpurun_out
mkdir -p $O
rm -f $O/parity_bench_shapes.jsonl
timeout 1500 python -m pytest tests -m gpu -q > $O/r2b_tests.log 2>&1
echo "tests rc=$?" >> $O/r2b_tests.log
./build/stream_ceiling 59336 > $O/r2b_ceiling.txt 2>&1
timeout 300 python bench.py --no-cpu --steps 1000 > $O/r2b_bench.json 2> $O/r2b_bench.err
timeout 300 python bench.py --no-cpu --steps 1000 --plan cached > $O/r2b_bench_cached.json 2> $O/r2b_bench_cached.err
timeout 300 python bench.py --no-cpu --steps 1000 --tune fuse_score=1 > $O/r2b_bench_fuse1.json 2> $O/r2b_bench_fuse1.err
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r2b_bench_driver.json 2> $O/r2b_bench_driver.err
python scripts/prof_step.py --reserve 1 > $O/r2b_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 150 -c 120 --csv --log-file $O/r2b_launches_warm.csv python scripts/prof_step.py --reserve 1 > $O/r2b_ncu.log 2>&1
echo done
