#!/bin/bash
# Round 2, GPU call L (8 GPUs): fused, three-deep pipelined mode B -- parity script and the scaling lines.
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514"
timeout 600 $TR scripts/modeb_p2p_check.py > $O/r2l_p2pcheck_n8.log 2>&1
echo "p2p rc=$?" >> $O/r2l_p2pcheck_n8.log
timeout 600 $TR bench.py --gpus 8 --steps 300 --warmup 10 > $O/r2l_bench_n8.json 2> $O/r2l_bench_n8.err
echo "rc=$?" >> $O/r2l_bench_n8.err
timeout 900 $TR bench.py --gpus 8 --steps 200 --warmup 10 --workload big100m > $O/r2l_bench_n8_big100m.json 2> $O/r2l_bench_n8_big100m.err
echo "rc=$?" >> $O/r2l_bench_n8_big100m.err
timeout 600 $TR bench.py --gpus 8 --steps 200 --warmup 10 --workload sideinfo > $O/r2l_bench_n8_sideinfo.json 2> $O/r2l_bench_n8_sideinfo.err
echo "rc=$?" >> $O/r2l_bench_n8_sideinfo.err
echo done
