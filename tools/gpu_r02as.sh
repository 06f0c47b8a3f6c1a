#!/bin/bash
O=gpurun_out
line() { python -c "import json; d=json.loads(open('$1').read().strip().splitlines()[-1]); r=d['roofline']; p=d['parity']; print('$2', round(d['value']/1e3,1), 'Gbit/s ms/step', round(d['ms_per_step'],4), 'kernel', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],4), 'parity', p['bit_mismatches'], p['beyond'])"; }
for v in b200 _chunks b200 _chunks; do
  export GF3_LIB_PATH=$PWD/gf3-audio-modem_b200/lib/libgf3$v.so
  for w in c3 w2048 c3-raw; do
    python bench.py --workload $w --steps 30 --no-cpu --no-e2e > $O/r02as_${w}$v.json 2> $O/r02as$v.err
    line $O/r02as_${w}$v.json "$v $w" || tail -c 300 $O/r02as$v.err
  done
done
unset GF3_LIB_PATH
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
