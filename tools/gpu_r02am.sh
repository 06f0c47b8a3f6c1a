#!/bin/bash
O=gpurun_out
for path in fused split; do
GF3_XCORR_PATH=$path timeout 300 python bench.py --workload c3-raw --steps 10 --no-cpu --no-e2e > $O/r02am_c3raw_$path.json 2> $O/r02am.err || tail -c 600 $O/r02am.err
python -c "import json; d=json.loads(open('$O/r02am_c3raw_$path.json').read().strip().splitlines()[-1]); r=d['roofline']; p=d['parity']; print('$path', round(d['value']/1e3,1),'Gbit/s', round(d['ms_per_step'],3),'ms', r['stages_ms'], 'parity', p['bit_mismatches'], p['beyond'], d['check']['streams_sync_failed'])"
done
GF3_XCORR_PATH=split timeout 600 python -m pytest tests/test_gpu_scale_parity.py -q -m gpu -p no:cacheprovider -x -k "sync" 2>&1 | tail -3
