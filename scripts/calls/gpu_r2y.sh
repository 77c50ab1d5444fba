#!/bin/bash
# Round 2, GPU call Y (1 GPU): graph overheads (plan interference vs launch gaps).
O=gpurun_out
mkdir -p $O
timeout 600 python scripts/graph_overheads.py ml20m > $O/r2y_graph_overheads.txt 2>&1
timeout 600 python scripts/graph_overheads.py sideinfo >> $O/r2y_graph_overheads.txt 2>&1
echo done
