#!/bin/bash
# Round 2, GPU call K (2 GPUs): three-deep pipeline of mode B.
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout 600 $TR scripts/modeb_p2p_check.py > $O/r2k_p2pcheck.log 2>&1
echo "p2p rc=$?" >> $O/r2k_p2pcheck.log
timeout 600 $TR bench.py --gpus 2 --steps 300 --warmup 10 > $O/r2k_bench_n2.json 2> $O/r2k_bench_n2.err
echo "rc=$?" >> $O/r2k_bench_n2.err
echo done
