#!/bin/bash
# C3 knob re-sweep on the final kernel: pilot loads in flight, phase-B unroll, flush unroll
O=gpurun_out
for rep in 1 2; do
for v in b200 _estu10 _estu40 _pbu4 _flu4; do
  export GF3_LIB_PATH=$PWD/gf3-audio-modem_b200/lib/libgf3$v.so
  python bench.py --steps 30 --no-cpu --no-e2e --no-parity > $O/r02ay_c3$v.json 2> $O/r02ay.err
  python -c "import json; d=json.loads(open('$O/r02ay_c3$v.json').read().strip().splitlines()[-1]); r=d['roofline']; print('$v', round(d['value']/1e3,1), 'kernel', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],4))" || tail -c 300 $O/r02ay.err
done
done
