#!/bin/bash
# usage: bash tools/run_variants.sh <variant> ...   (libs built as gf3-audio-modem_b200/lib/libgf3b200_<variant>.so,
# e.g. GF3_LIB_NAME=libgf3b200_u8.so GF3_EXTRA_FLAGS="-DGF3_PHASEB_UNROLL=8" python gf3-audio-modem_b200/build.py)
# prints, per variant, a parity smoke result and the receive-chain time (tools/bench_chain.py)
for v in "$@"; do
  L=$PWD/gf3-audio-modem_b200/lib/libgf3b200_$v.so
  [ "$v" = default ] && L=$PWD/gf3-audio-modem_b200/lib/libgf3b200.so
  GF3_LIB_PATH=$L python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stage_receive or loopback" 2>&1 | tail -1
  for w in ${WORKLOADS:-c3}; do
    echo "variant $v $w: $(GF3_LIB_PATH=$L python tools/bench_chain.py $w 40 | tail -2 | tr '\n' ' ')"
  done
done
