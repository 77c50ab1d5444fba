#!/bin/bash
# Round 2, GPU call J (2 GPUs): owner update without the separate ordered-sum pass; real-rank checks + bench.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dp.py tests/test_gpu_multi.py -m gpu -q -x > $O/r2j_tests.log 2>&1
echo "tests rc=$?" >> $O/r2j_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR bench.py --gpus 2 --steps 300 --warmup 10 > $O/r2j_bench_n2.json 2> $O/r2j_bench_n2.err
echo "rc=$?" >> $O/r2j_bench_n2.err
timeout 600 $TR bench.py --gpus 2 --steps 300 --warmup 10 --slack 0.55 > $O/r2j_bench_n2_slack55.json 2> $O/r2j_bench_n2_slack55.err
echo "rc=$?" >> $O/r2j_bench_n2_slack55.err
echo done
