#!/bin/bash
# Round 2, GPU call F: pipe kernel v2 (gradient row prefetched a round ahead, bias loads under the rounds), graph knobs.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sampled.py tests/test_gpu_dp.py tests/test_gpu_closed.py -m gpu -q -x > $O/r2f_tests.log 2>&1
echo "tests rc=$?" >> $O/r2f_tests.log
timeout 300 python bench.py --no-cpu --steps 1000 > $O/r2f_bench.json 2> $O/r2f_bench.err
timeout 300 python bench.py --no-cpu --steps 1000 --reserve 0 > $O/r2f_bench_res0.json 2> $O/r2f_bench_res0.err
timeout 300 python bench.py --no-cpu --steps 1000 --side-priority 0 > $O/r2f_bench_prio0.json 2> $O/r2f_bench_prio0.err
timeout 300 python bench.py --no-cpu --steps 1000 --reserve 0 --side-priority 0 > $O/r2f_bench_res0_prio0.json 2> $O/r2f_bench_res0_prio0.err
timeout 300 python bench.py --no-cpu --steps 1000 --tune adam_reserve=1 > $O/r2f_bench_ares.json 2> $O/r2f_bench_ares.err
timeout 300 python bench.py --no-cpu --steps 1000 --plan cached > $O/r2f_bench_cached.json 2> $O/r2f_bench_cached.err
python scripts/prof_step.py --reserve 1 > $O/r2f_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 78 -c 52 --csv --log-file $O/r2f_launches_warm.csv python scripts/prof_step.py --reserve 1 > $O/r2f_ncu.log 2>&1
echo done
