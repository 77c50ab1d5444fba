#!/bin/bash
# Round 2, GPU call AU (2 GPUs): mode A (replicated tables, dense all-reduce) with real-rank parity in the line.
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR bench.py --gpus 2 --steps 50 --warmup 5 --parallel dp > $O/r2au_bench_n2_modeA.json 2> $O/r2au_bench_n2_modeA.err
timeout 600 $TR bench.py --gpus 2 --steps 30 --warmup 3 --parallel dp --workload sideinfo > $O/r2au_bench_n2_modeA_sideinfo.json 2> $O/r2au_bench_n2_modeA_sideinfo.err
echo done
