#!/bin/bash
# Round 2, GPU call AZ (1 GPU): smoke + suite on the library rebuilt from the final commit.
O=gpurun_out
mkdir -p $O
python -c "from vae_b200 import _lib; print('stale', _lib._stale())" > $O/r2az_stale.txt 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2az_smoke.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > $O/r2az_tests.log 2>&1
echo "tests rc=$?" >> $O/r2az_tests.log
echo done
