#!/bin/bash
# Round 2, GPU call AY (1 GPU): where the plan branch forks from the step (start / after k_stage / after k_score).
# (needed an experiment knob that was not kept: ml20m 102.3 / 103.0 / 110.1 us, sideinfo 263.1 / 255.1 us)
O=gpurun_out
mkdir -p $O
for rep in 1 2; do
for fa in 0 1 3; do
  VFMB_FORK_AFTER=$fa timeout 300 python bench.py --no-cpu --steps 1000 > $O/r2ay_fork${fa}_$rep.json 2> $O/r2ay_fork${fa}_$rep.err
done
done
VFMB_FORK_AFTER=1 timeout 300 python bench.py --no-cpu --steps 300 --workload sideinfo > $O/r2ay_side_fork1.json 2> $O/r2ay_side_fork1.err
VFMB_FORK_AFTER=0 timeout 300 python bench.py --no-cpu --steps 300 --workload sideinfo > $O/r2ay_side_fork0.json 2> $O/r2ay_side_fork0.err
echo done
