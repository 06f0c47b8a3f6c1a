#!/bin/bash
O=gpurun_out
python bench.py --workload a2-raw --steps 10 --no-cpu > $O/r02z_a2raw.json 2> $O/r02z.err || tail -c 600 $O/r02z.err
python -c "import json; d=json.loads(open('$O/r02z_a2raw.json').read().strip().splitlines()[-1]); r=d['roofline']; p=d['parity']; print(round(d['value']/1e3,1),'Gbit/s', round(d['ms_per_step'],3),'ms', r['stages_ms'], 'sync', r['sync'], 'parity', p['bit_mismatches'], p['beyond'], 'e2e', d['e2e']['value'], d['check'])"
