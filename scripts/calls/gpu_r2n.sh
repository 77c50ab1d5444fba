#!/bin/bash
# Round 2, GPU call N (2 GPUs): graph knobs of the pipelined mode B; sideinfo in mode A vs mode B.
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515"
for cfg in "1 -1" "1 0" "0 -1" "2 -1"; do
  set -- $cfg
  timeout 300 $TR bench.py --gpus 2 --steps 300 --warmup 10 --reserve $1 --side-priority $2 > $O/r2n_bench_n2_res$1_prio$2.json 2> $O/r2n_bench_n2_res$1_prio$2.err
done
timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 5 --workload sideinfo > $O/r2n_bench_n2_sideinfo_modeB.json 2> $O/r2n_bench_n2_sideinfo_modeB.err
timeout 600 $TR bench.py --gpus 2 --steps 30 --warmup 3 --workload sideinfo --parallel dp > $O/r2n_bench_n2_sideinfo_modeA.json 2> $O/r2n_bench_n2_sideinfo_modeA.err
timeout 900 $TR bench.py --gpus 2 --steps 100 --warmup 5 --workload big100m > $O/r2n_bench_n2_big100m.json 2> $O/r2n_bench_n2_big100m.err
echo done
