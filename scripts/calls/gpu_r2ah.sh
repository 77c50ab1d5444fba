#!/bin/bash
# Round 2, GPU call AH (4 GPUs): mode B at N = 4 and the peer-memory cross-check with the final kernels.
O=gpurun_out
mkdir -p $O
python -c "from vae_b200 import _lib; print('stale', _lib._stale())" > $O/r2ah_stale.txt 2>&1
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523"
timeout 300 $TR2 scripts/modeb_p2p_check.py > $O/r2ah_p2p_check_n2.txt 2>&1
timeout 300 $TR scripts/modeb_p2p_check.py > $O/r2ah_p2p_check_n4.txt 2>&1
timeout 300 $TR bench.py --gpus 4 --steps 300 --warmup 10 > $O/r2ah_bench_n4.json 2> $O/r2ah_bench_n4.err
timeout 600 $TR bench.py --gpus 4 --steps 100 --warmup 5 --workload big100m > $O/r2ah_bench_n4_big100m.json 2> $O/r2ah_bench_n4_big100m.err
timeout 600 $TR bench.py --gpus 4 --steps 100 --warmup 5 --workload sideinfo > $O/r2ah_bench_n4_sideinfo.json 2> $O/r2ah_bench_n4_sideinfo.err
echo done
