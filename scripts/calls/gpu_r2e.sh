#!/bin/bash
# Round 2, GPU call E: ncu --set full of the step kernels (warm caches); only the CSV pages travel back.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sampled.py tests/test_gpu_dp.py tests/test_gpu_bench_shapes.py -m gpu -q -x > $O/r2e_tests.log 2>&1
echo "tests rc=$?" >> $O/r2e_tests.log
timeout 300 python bench.py --no-cpu --steps 1000 > $O/r2e_bench.json 2> $O/r2e_bench.err
timeout 300 python bench.py --no-cpu --steps 1000 --tune prefetch_mv=3 > $O/r2e_bench_pf3.json 2> $O/r2e_bench_pf3.err
python scripts/prof_step.py --reserve 1 > $O/r2e_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k_adam_rows_pipe|k_score|k_gather|k_stage' -s 16 -c 4 -o /tmp/prof_r02e python scripts/prof_step.py --reserve 1 > $O/r2e_ncu.log 2>&1
ncu -i /tmp/prof_r02e.ncu-rep --page raw --csv > $O/r2e_ncu_raw.csv 2>/dev/null
ncu -i /tmp/prof_r02e.ncu-rep --page source --csv > $O/r2e_ncu_source.csv 2>/dev/null
ls -la /tmp/prof_r02e.ncu-rep >> $O/r2e_ncu.log
sz=$(stat -c %s /tmp/prof_r02e.ncu-rep); if [ "$sz" -lt 45000000 ]; then cp /tmp/prof_r02e.ncu-rep $O/; fi
du -sh $O >> $O/r2e_ncu.log
echo done
