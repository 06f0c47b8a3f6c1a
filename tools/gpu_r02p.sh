#!/bin/bash
O=gpurun_out
for v in b200 _alt12 _alt12f; do
  export GF3_LIB_PATH=$PWD/gf3-audio-modem_b200/lib/libgf3$v.so
  for w in c4 a2 c4-long; do
    python bench.py --workload $w --steps 30 --no-cpu --no-e2e > $O/r02p_${w}$v.json 2> $O/r02p$v.err
    python -c "import json; d=json.loads(open('$O/r02p_${w}$v.json').read().strip().splitlines()[-1]); r=d['roofline']; p=d['parity']; print('$v $w', round(d['value']/1e3,1), 'Gbit/s ms/step', round(d['ms_per_step'],4), 'kernel', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],4), 'chain', round(r['chain']['frac'],4), 'parity', p['bit_mismatches'], p['beyond'], p.get('within_eq_tol'))" || tail -c 300 $O/r02p$v.err
  done
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -q -m gpu -p no:cacheprovider -k "4096 or kat1 or edge or roundtrip or staged" 2>&1 | tail -2
done
