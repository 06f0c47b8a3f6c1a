#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_scale_parity.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "long_chirp or a2 or kat1 or kat4 or sync or xcorr" 2>&1 | tail -8
for m in detect dense; do
extra=""; [ $m = dense ] && extra="--dense-sync"
timeout 300 python bench.py --workload a2-raw --steps 10 --no-cpu --no-e2e $extra > $O/r02aj_a2raw_$m.json 2> $O/r02aj.err || tail -c 600 $O/r02aj.err
python -c "import json; d=json.loads(open('$O/r02aj_a2raw_$m.json').read().strip().splitlines()[-1]); r=d['roofline']; p=d['parity']; print('$m', round(d['value']/1e3,1),'Gbit/s', round(d['ms_per_step'],3),'ms', r['stages_ms'], 'parity', p['bit_mismatches'], p['beyond'], d['check']['streams_sync_failed'])"
done
timeout 300 python bench.py --workload a2-raw --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/r02aj_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02aj_launches_a2raw.csv python bench.py --workload a2-raw --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/r02aj_ncu1.log 2>&1
grep -c xcorr $O/r02aj_launches_a2raw.csv
