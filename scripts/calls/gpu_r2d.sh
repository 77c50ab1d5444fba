#!/bin/bash
# Round 2, GPU call D: cp.async-pipelined row update (k_adam_rows_pipe) vs the register-staged kernel.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sampled.py tests/test_gpu_closed.py tests/test_gpu_dp.py -m gpu -q -x > $O/r2d_tests.log 2>&1
echo "tests rc=$?" >> $O/r2d_tests.log
python scripts/adam_micro.py > $O/r2d_adam_micro.txt 2>&1
timeout 300 python bench.py --no-cpu --steps 1000 > $O/r2d_bench.json 2> $O/r2d_bench.err
timeout 300 python bench.py --no-cpu --steps 1000 --tune adam_pipe=0 > $O/r2d_bench_nopipe.json 2> $O/r2d_bench_nopipe.err
timeout 300 python bench.py --no-cpu --steps 1000 --plan cached > $O/r2d_bench_cached.json 2> $O/r2d_bench_cached.err
timeout 300 python bench.py --no-cpu --steps 300 --workload sideinfo > $O/r2d_bench_sideinfo.json 2> $O/r2d_bench_sideinfo.err
python scripts/prof_step.py --reserve 1 > $O/r2d_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 78 -c 52 --csv --log-file $O/r2d_launches_warm.csv python scripts/prof_step.py --reserve 1 > $O/r2d_ncu.log 2>&1
echo done
