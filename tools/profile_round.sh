#!/bin/bash
# Round-end evidence run (1 GPU): bench lines, ncu launch list, ncu --set full of the top kernel.
# usage: bash tools/profile_round.sh <tag>     (outputs under gpurun_out/)
TAG=${1:-rXX}
set -x
python bench.py > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
python bench.py --workload c4 --no-cpu > gpurun_out/${TAG}_bench_c4.json 2>> gpurun_out/${TAG}_bench_c3.err
python bench.py --workload a2 --no-cpu > gpurun_out/${TAG}_bench_a2.json 2>> gpurun_out/${TAG}_bench_c3.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench_c3.err
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rx_demod|rx_estimate" -s 3 -c 1 -o gpurun_out/${TAG}_prof_c3 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rx_demod" -s 3 -c 1 -o gpurun_out/${TAG}_prof_c4 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu3.log 2>&1
tail -2 gpurun_out/${TAG}_ncu3.log
cat gpurun_out/${TAG}_bench_c3.json
