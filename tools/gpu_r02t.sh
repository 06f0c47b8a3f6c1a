#!/bin/bash
O=gpurun_out
for v in _alt12c; do
  export GF3_LIB_PATH=$PWD/gf3-audio-modem_b200/lib/libgf3$v.so
  for w in c4 a2; do
    python bench.py --workload $w --steps 30 --no-cpu --no-e2e > $O/r02t_${w}$v.json 2> $O/r02t$v.err
    python -c "import json; d=json.loads(open('$O/r02t_${w}$v.json').read().strip().splitlines()[-1]); r=d['roofline']; p=d['parity']; print('$v $w', round(d['value']/1e3,1), 'Gbit/s ms/step', round(d['ms_per_step'],4), 'kernel', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],4), 'chain', round(r['chain']['frac'],4), 'parity', p['bit_mismatches'], p['beyond'], p.get('within_eq_tol'))" || tail -c 300 $O/r02t$v.err
  done
  GF3_RX_STAGED=1 python bench.py --workload c4 --steps 30 --no-cpu --no-e2e > $O/r02t_c4_staged$v.json 2> $O/r02t$v.err
  python -c "import json; d=json.loads(open('$O/r02t_c4_staged$v.json').read().strip().splitlines()[-1]); r=d['roofline']; print('$v c4 staged', round(d['ms_per_step'],4), 'kernel', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],4))" || tail -c 300 $O/r02t$v.err
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -q -m gpu -p no:cacheprovider -k "4096 or kat1 or edge" 2>&1 | tail -2
done
