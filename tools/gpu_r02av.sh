#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale_parity.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "long_chirp or mid_length or a2 or kat or sync or xcorr or raw or streams" 2>&1 | tail -3
timeout 300 python bench.py --workload a2-raw --steps 20 --no-cpu --no-e2e > $O/r02av_a2raw.json 2> $O/r02av.err || tail -c 600 $O/r02av.err
python -c "import json; d=json.loads(open('$O/r02av_a2raw.json').read().strip().splitlines()[-1]); r=d['roofline']; p=d['parity']; print(round(d['value']/1e3,1),'Gbit/s', round(d['ms_per_step'],3),'ms', r['stages_ms'], 'parity', p['bit_mismatches'], p['beyond'], d['check']['streams_sync_failed'])"
timeout 300 python bench.py --workload a2-raw --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/r02av_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02av_launches_a2raw.csv python bench.py --workload a2-raw --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/r02av_ncu1.log 2>&1
grep -c xcorr $O/r02av_launches_a2raw.csv
