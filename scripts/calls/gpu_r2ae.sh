#!/bin/bash
# Round 2, GPU call AE (1 GPU): short timed regions (driver arguments) after moving the NVML sampler start.
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_sampled.py -m gpu -q -x -k "knobs" > $O/r2ae_tests.log 2>&1
for rep in 1 2 3; do
  timeout 300 python bench.py --no-cpu --steps 20 --warmup 5 > $O/r2ae_bench_k20_$rep.json 2> $O/r2ae_bench_k20_$rep.err
done
timeout 300 python bench.py --no-cpu --steps 200 --warmup 5 > $O/r2ae_bench_k200.json 2> $O/r2ae_bench_k200.err
timeout 300 python bench.py --no-cpu > $O/r2ae_bench_default.json 2> $O/r2ae_bench_default.err
for w in fraction ml100k; do
  timeout 300 python bench.py --no-cpu --steps 300 --workload $w > $O/r2ae_bench_$w.json 2> $O/r2ae_bench_$w.err
  timeout 300 python bench.py --no-cpu --steps 300 --workload $w --tune gather_wide=0 > $O/r2ae_bench_${w}_gw0.json 2> $O/r2ae_bench_${w}_gw0.err
done
python scripts/prof_step.py --reserve 1 --workload fraction --rows 10720 > $O/r2ae_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 78 -c 52 --csv --log-file $O/r2ae_launches_fraction.csv python scripts/prof_step.py --reserve 1 --workload fraction --rows 10720 > $O/r2ae_ncu.log 2>&1
echo done
