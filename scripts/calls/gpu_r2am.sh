#!/bin/bash
# Round 2, GPU call AM (1 GPU): units per warp pass in k_stage / k_score (16 vs 32).
O=gpurun_out
mkdir -p $O
run() { name=$1; shift; timeout 300 "$@" > $O/r2am_$name.json 2> $O/r2am_$name.err; }
run base python bench.py --no-cpu --steps 1000
run st16 python bench.py --no-cpu --steps 1000 --tune stage_chunk=16
run sc16 python bench.py --no-cpu --steps 1000 --tune score_chunk=16
run both16 python bench.py --no-cpu --steps 1000 --tune stage_chunk=16 --tune score_chunk=16
run both16_sw1 python bench.py --no-cpu --steps 1000 --tune stage_chunk=16 --tune score_chunk=16 --tune score_wide=1 --tune stage_wide=1
run side_base python bench.py --no-cpu --steps 300 --workload sideinfo
run side_both16 python bench.py --no-cpu --steps 300 --workload sideinfo --tune stage_chunk=16 --tune score_chunk=16
python scripts/tune_variants.py stage_chunk=32,score_chunk=32 stage_chunk=16,score_chunk=16 > $O/r2am_variants_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"k_score|k_gather|k_stage|k_adam" --csv \
    --log-file $O/r2am_variants.csv python scripts/tune_variants.py stage_chunk=32,score_chunk=32 stage_chunk=16,score_chunk=16 > $O/r2am_ncu.log 2>&1
echo done
