#!/bin/bash
# usage: bash tools/run_variants.sh <variant> ...   (libs built as gf3-audio-modem_b200/lib/libgf3b200_<variant>.so)
# prints, per variant, a parity smoke result and the C3 demod kernel time
for v in "$@"; do
  L=$PWD/gf3-audio-modem_b200/lib/libgf3b200_$v.so
  [ "$v" = default ] && L=$PWD/gf3-audio-modem_b200/lib/libgf3b200.so
  GF3_LIB_PATH=$L python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stage_receive or loopback" 2>&1 | tail -1
  for w in ${WORKLOADS:-c3}; do
  GF3_LIB_PATH=$L python bench.py --workload $w --no-cpu --no-e2e --steps 30 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('variant','$v','$w','ms/step',round(d['ms_per_step'],4),'demod ms',round(r['avg_launch_ms'],4),'frac',round(r['frac'],4))"
  done
done
