#!/bin/bash
# Round-end evidence run (1 GPU): bench lines of every workload, ncu launch lists, ncu --set full of the top kernels.
# usage: bash tools/profile_round.sh <tag>     (outputs under gpurun_out/; copy what is judged into profiles/)
TAG=${1:-rXX}
O=gpurun_out
set -x
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -3
python bench.py > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench.err
python bench.py --workload c3-raw --steps 20 > $O/${TAG}_bench_c3raw.json 2>> $O/${TAG}_bench.err
python bench.py --workload a2-raw --steps 20 > $O/${TAG}_bench_a2raw.json 2>> $O/${TAG}_bench.err
python bench.py --workload c4 --no-cpu > $O/${TAG}_bench_c4.json 2>> $O/${TAG}_bench.err
python bench.py --workload c4-long --no-cpu > $O/${TAG}_bench_c4long.json 2>> $O/${TAG}_bench.err
python bench.py --workload a2 --no-cpu > $O/${TAG}_bench_a2.json 2>> $O/${TAG}_bench.err
python bench.py --workload w2048 --no-cpu > $O/${TAG}_bench_w2048.json 2>> $O/${TAG}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err
python bench.py --impl reference --workload c3-raw --steps 3 --warmup 1 > $O/${TAG}_bench_ref_c3raw.json 2>> $O/${TAG}_bench.err
python tools/bench_stages.py > $O/${TAG}_stages.md 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/${TAG}_ncu1.log 2>&1
python bench.py --workload c3-raw --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/${TAG}_plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches_c3raw.csv python bench.py --workload c3-raw --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/${TAG}_ncu1b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rx_demod" -s 3 -c 1 -o $O/${TAG}_prof_c3 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rx_demod" -s 3 -c 1 -o $O/${TAG}_prof_c4 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/${TAG}_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"xcorr_fused" -s 2 -c 1 -o $O/${TAG}_prof_xcorr python bench.py --workload c3-raw --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/${TAG}_ncu4.log 2>&1
tail -2 $O/${TAG}_ncu4.log
head -c 1500 $O/${TAG}_bench_c3.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
