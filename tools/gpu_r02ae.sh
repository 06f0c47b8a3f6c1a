#!/bin/bash
O=gpurun_out
for mb in 16 32 48 64 96 2048; do
GF3_XC_TILE_MB=$mb python bench.py --workload a2-raw --steps 10 --no-cpu --no-e2e --no-parity > $O/r02ae_a2raw_$mb.json 2> $O/r02ae.err || tail -c 600 $O/r02ae.err
python -c "import json; d=json.loads(open('$O/r02ae_a2raw_$mb.json').read().strip().splitlines()[-1]); r=d['roofline']; print('tile MB=$mb', round(d['value']/1e3,1),'Gbit/s', round(d['ms_per_step'],3),'ms', r['stages_ms'])"
done
