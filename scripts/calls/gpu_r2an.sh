#!/bin/bash
# Round 2, GPU call AN (1 GPU): ncu --set full with source counters of k_stage (warm step).
O=gpurun_out
mkdir -p $O
python scripts/prof_step.py --reserve 1 > $O/r2an_prof_plain.log 2>&1 &&
timeout 600 ncu --set full --import-source on --clock-control none --cache-control none -k regex:k_stage -s 6 -c 1 -f -o /tmp/r2an \
    python scripts/prof_step.py --reserve 1 > $O/r2an_ncu.log 2>&1
ncu -i /tmp/r2an.ncu-rep --page raw --csv > $O/r2an_stage_raw.csv 2>/dev/null
ncu -i /tmp/r2an.ncu-rep --page source --csv > $O/r2an_stage_source.csv 2>/dev/null
echo done
