#!/bin/bash
O=gpurun_out
for v in b200 _notwab; do
  GF3_LIB_PATH=$PWD/gf3-audio-modem_b200/lib/libgf3$v.so python bench.py --steps 40 --no-cpu --no-e2e > $O/r02m_c3$v.json 2> $O/r02m$v.err
  python -c "import json; d=json.loads(open('$O/r02m_c3$v.json').read().strip().splitlines()[-1]); print('$v', round(d['value']/1e3,1), 'Gbit/s', round(d['roofline']['avg_launch_ms'],4), 'ms', round(d['roofline']['frac'],4), d['parity']['bit_mismatches'], d['parity']['beyond'])" || tail -c 300 $O/r02m$v.err
done
python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
