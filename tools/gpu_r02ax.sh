#!/bin/bash
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tx_symbols -c 4 -o $O/r02ax_prof_tx python tools/bench_stages.py > $O/r02ax_ncu.log 2>&1
tail -1 $O/r02ax_ncu.log
