#!/bin/bash
# final numbers of the long-chirp matched filter (a2-raw), the ncu evidence, and the whole GPU suite on the final tree
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
timeout 300 python bench.py --workload a2-raw --steps 20 > $O/r02an_bench_a2raw.json 2> $O/r02an.err || tail -c 600 $O/r02an.err
timeout 300 python bench.py --workload a2-raw --steps 20 --dense-sync --no-cpu --no-e2e > $O/r02an_bench_a2raw_dense.json 2>> $O/r02an.err
GF3_XCORR_MAC=0 timeout 300 python bench.py --workload a2-raw --steps 20 --dense-sync --no-cpu --no-e2e > $O/r02an_bench_a2raw_dense_r01kernels.json 2>> $O/r02an.err
timeout 300 python bench.py --impl reference --workload a2-raw --steps 2 --warmup 1 > $O/r02an_bench_ref_a2raw.json 2>> $O/r02an.err
for f in a2raw a2raw_dense a2raw_dense_r01kernels ref_a2raw; do python -c "import json; d=json.loads(open('$O/r02an_bench_$f.json').read().strip().splitlines()[-1]); r=d.get('roofline') or {}; print('$f', round(d['value']/1e3,2),'Gbit/s', round(d['ms_per_step'],3),'ms', r.get('stages_ms'), (d.get('parity') or {}).get('bit_mismatches'), (d.get('e2e') or {}).get('value'))"; done
timeout 300 python bench.py --workload a2-raw --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/r02an_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02an_launches_a2raw.csv python bench.py --workload a2-raw --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/r02an_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"xcorr_fwd|xcorr_bound|xcorr_acc" -s 4 -c 4 -o $O/r02an_prof_a2raw python bench.py --workload a2-raw --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/r02an_ncu2.log 2>&1
tail -1 $O/r02an_ncu2.log
