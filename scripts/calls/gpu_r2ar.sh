#!/bin/bash
# Round 2, GPU call AR (2 GPUs): real-rank cross-check and bench line with the final library.
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
python -c "from vae_b200 import _lib; print('stale', _lib._stale())" > $O/r2ar_stale.txt 2>&1
timeout 300 $TR scripts/modeb_p2p_check.py > $O/r2ar_p2p_check_n2.txt 2>&1
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2ar_bench_n2_k20.json 2> $O/r2ar_bench_n2_k20.err
timeout 300 $TR bench.py --gpus 2 --steps 300 --warmup 10 > $O/r2ar_bench_n2.json 2> $O/r2ar_bench_n2.err
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > $O/r2ar_bench_n2_reference.json 2> $O/r2ar_bench_n2_reference.err
echo done
