#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale_parity.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "sync or xcorr or raw or streams or kat" 2>&1 | tail -3
for w in c3-raw; do
timeout 300 python bench.py --workload $w --steps 20 --no-cpu --no-e2e > $O/r02ap_$w.json 2> $O/r02ap.err || tail -c 600 $O/r02ap.err
python -c "import json; d=json.loads(open('$O/r02ap_$w.json').read().strip().splitlines()[-1]); r=d['roofline']; p=d['parity']; print('$w', round(d['value']/1e3,1),'Gbit/s', round(d['ms_per_step'],3),'ms', r['stages_ms'], 'parity', p['bit_mismatches'], p['beyond'], d['check']['streams_sync_failed'])"
done
timeout 300 python tools/debug_sparse.py 2>&1 | tail -5
