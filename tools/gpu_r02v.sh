#!/bin/bash
# float32 packets at odd sample offsets: aligned loads + lane shuffle (default build) against scalar loads (_noshfl)
O=gpurun_out
line() { python -c "import json; d=json.loads(open('$1').read().strip().splitlines()[-1]); r=d['roofline']; p=d['parity']; print('$2', round(d['value']/1e3,1), 'Gbit/s ms/step', round(d['ms_per_step'],4), 'kernel', r.get('kernel','')[:24], round(r['avg_launch_ms'],4), 'frac', round(r['frac'],4), 'stages', r.get('stages_ms'), 'parity', p['bit_mismatches'], p['beyond'], p.get('within_eq_tol'))"; }
for v in b200 _noshfl b200 _noshfl; do
  export GF3_LIB_PATH=$PWD/gf3-audio-modem_b200/lib/libgf3$v.so
  for w in c3 c3-raw w2048; do
    python bench.py --workload $w --steps 30 --no-cpu --no-e2e > $O/r02v_${w}$v.json 2> $O/r02v$v.err
    line $O/r02v_${w}$v.json "$v $w" || tail -c 300 $O/r02v$v.err
  done
done
unset GF3_LIB_PATH
python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
python tools/bench_stages.py 2>&1 | tail -12
