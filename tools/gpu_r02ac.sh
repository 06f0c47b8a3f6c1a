#!/bin/bash
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"xcorr_mac" -s 1 -c 1 -o $O/r02ac_prof_a2raw python bench.py --workload a2-raw --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/r02ac_ncu.log 2>&1
tail -1 $O/r02ac_ncu.log
