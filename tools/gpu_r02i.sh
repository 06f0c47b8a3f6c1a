#!/bin/bash
# Round 2, GPU call I: partial register prefetch variants of the C3 kernel.
O=gpurun_out
for v in b200 _pf4 _pf8 _pf16; do
  lib=gf3-audio-modem_b200/lib/libgf3$v.so
  GF3_LIB_PATH=$PWD/$lib python bench.py --steps 40 --no-cpu --no-e2e --no-parity > $O/r02i_c3$v.json 2> $O/r02i$v.err
  python -c "import json; d=json.loads(open('$O/r02i_c3$v.json').read().strip().splitlines()[-1]); print('$v', round(d['value']/1e3,1), 'Gbit/s', round(d['roofline']['avg_launch_ms'],4), 'ms', round(d['roofline']['frac'],4), d['check']['ber'])" || tail -c 300 $O/r02i$v.err
done
