#!/bin/bash
# Round 2, GPU call C: fused matched filter variants (CTAs per SM) + the at-scale parity tests.
O=gpurun_out
( time python -m pytest tests/test_gpu_scale_parity.py tests/test_gpu_parity.py -q -m gpu -rA -p no:cacheprovider -k "sync or c3_fused or quirk or peak" ) > $O/r02c_pytest.log 2>&1
tail -8 $O/r02c_pytest.log
for mb in 2 3; do
  GF3_XC_MINB=$mb python bench.py --workload c3-raw --steps 10 --no-cpu --no-e2e --no-parity --split-sync > $O/r02c_c3raw_minb$mb.json 2> $O/r02c_minb$mb.err; tail -c 300 $O/r02c_minb$mb.err
done
GF3_XC_MINB=2 ncu --set full --clock-control none --import-source on -k regex:"xcorr_fused" -s 2 -c 1 -o $O/r02c_prof_xcorr_minb2 python bench.py --workload c3-raw --streams 1024 --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/r02c_ncu.log 2>&1
python - <<'PY'
import json
for f in ["r02c_c3raw_minb2.json","r02c_c3raw_minb3.json"]:
    try:
        d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
        print(f, "%.1f Gbit/s" % (d["value"]/1e3), d["roofline"]["stages_ms"], "frac", d["roofline"]["frac"])
    except Exception as e:
        print(f, "ERR", e)
PY
