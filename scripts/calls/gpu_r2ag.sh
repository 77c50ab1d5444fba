#!/bin/bash
# Round 2, GPU call AG (8 GPUs): mode B at N = 8 with the final kernels (ml20m, big100m, sideinfo) + p2p check.
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
timeout 300 $TR scripts/modeb_p2p_check.py > $O/r2ag_p2p_check_n8.txt 2>&1
timeout 300 $TR bench.py --gpus 8 --steps 300 --warmup 10 > $O/r2ag_bench_n8.json 2> $O/r2ag_bench_n8.err
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2ag_bench_n8_k20.json 2> $O/r2ag_bench_n8_k20.err
timeout 600 $TR bench.py --gpus 8 --steps 100 --warmup 5 --workload big100m > $O/r2ag_bench_n8_big100m.json 2> $O/r2ag_bench_n8_big100m.err
timeout 600 $TR bench.py --gpus 8 --steps 100 --warmup 5 --workload sideinfo > $O/r2ag_bench_n8_sideinfo.json 2> $O/r2ag_bench_n8_sideinfo.err
echo done
