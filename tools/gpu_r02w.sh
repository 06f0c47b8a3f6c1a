#!/bin/bash
O=gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -4
python tools/bench_stages.py > $O/r02w_stages.md 2>&1; tail -12 $O/r02w_stages.md
ncu --set full --clock-control none --import-source on -k regex:"tx_symbols" -s 2 -c 1 -o $O/r02w_prof_tx python tools/bench_stages.py > $O/r02w_ncu.log 2>&1
tail -1 $O/r02w_ncu.log
