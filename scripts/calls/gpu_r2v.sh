#!/bin/bash
# Round 2, GPU call V (1 GPU): where the step's time outside its four kernels goes (plan overlap, graph launch).
O=gpurun_out
mkdir -p $O
for cfg in "graph" "cached" "prefetch"; do
  timeout 300 python bench.py --no-cpu --steps 1000 --plan $cfg --tune score_wide=1 > $O/r2v_bench_$cfg.json 2> $O/r2v_bench_$cfg.err
done
timeout 300 python bench.py --no-cpu --steps 1000 --reserve 0 --tune score_wide=1 > $O/r2v_bench_reserve0.json 2> $O/r2v_bench_reserve0.err
timeout 300 python bench.py --no-cpu --steps 1000 --tune score_wide=1 --tune adam_reserve=1 > $O/r2v_bench_adamres.json 2> $O/r2v_bench_adamres.err
echo done
