#!/bin/bash
# Round 2, GPU call Z (2 GPUs): mode B with the restructured gather / score / stage kernels.
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q -x > $O/r2z_tests.log 2>&1
echo "tests rc=$?" >> $O/r2z_tests.log
timeout 300 $TR scripts/modeb_p2p_check.py > $O/r2z_p2p_check.txt 2>&1
timeout 300 $TR bench.py --gpus 2 --steps 300 --warmup 10 > $O/r2z_bench_n2.json 2> $O/r2z_bench_n2.err
timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 5 --workload sideinfo > $O/r2z_bench_n2_sideinfo.json 2> $O/r2z_bench_n2_sideinfo.err
timeout 900 $TR bench.py --gpus 2 --steps 100 --warmup 5 --workload big100m > $O/r2z_bench_n2_big100m.json 2> $O/r2z_bench_n2_big100m.err
echo done
