#!/bin/bash
# Round 2, GPU call W (1 GPU): step next to the concurrently built plan -- which resources to leave free.
O=gpurun_out
mkdir -p $O
run() { # name, env..., args
  name=$1; shift
  timeout 300 env "$@" > $O/r2w_$name.json 2> $O/r2w_$name.err
}
for rep in 1 2; do
  run base_$rep python bench.py --no-cpu --steps 1000 --tune score_wide=1
  run adamres_$rep python bench.py --no-cpu --steps 1000 --tune score_wide=1 --tune adam_reserve=1
  run g96_$rep VFMB_VARIANT=g96 VFMB_NVCC_EXTRA=-DVFMB_GATHER_MAXREG=96 python bench.py --no-cpu --steps 1000 --tune score_wide=1
  run g96_adamres_$rep VFMB_VARIANT=g96 VFMB_NVCC_EXTRA=-DVFMB_GATHER_MAXREG=96 python bench.py --no-cpu --steps 1000 --tune score_wide=1 --tune adam_reserve=1
done
run side_adamres VFMB_VARIANT=g96 VFMB_NVCC_EXTRA=-DVFMB_GATHER_MAXREG=96 python bench.py --no-cpu --steps 300 --workload sideinfo --tune score_wide=1 --tune adam_reserve=1
run side_base python bench.py --no-cpu --steps 300 --workload sideinfo --tune score_wide=1
echo done
