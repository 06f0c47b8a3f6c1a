#!/bin/bash
O=gpurun_out
for rep in 1 2; do
for v in b200 _flu4; do
  export GF3_LIB_PATH=$PWD/gf3-audio-modem_b200/lib/libgf3$v.so
  for w in c4 a2 w2048 c3-raw; do
  python bench.py --workload $w --steps 30 --no-cpu --no-e2e --no-parity > $O/r02az_${w}$v.json 2> $O/r02az.err
  python -c "import json; d=json.loads(open('$O/r02az_${w}$v.json').read().strip().splitlines()[-1]); r=d['roofline']; print('$v $w', round(d['value']/1e3,1), 'ms', round(d['ms_per_step'],4))" || tail -c 300 $O/r02az.err
  done
done
done
