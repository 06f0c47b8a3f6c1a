#!/bin/bash
# Round 2, GPU call B: fused matched filter -- parity, timing against the two-kernel form, ncu capture.
O=gpurun_out
mkdir -p $O
( time python -m pytest tests/test_gpu_scale_parity.py tests/test_gpu_parity.py tests/test_gpu_edges.py -q -m gpu -rA -p no:cacheprovider -k "sync or c3_fused or quirk or peak or roundtrip or sweep or stage_sync" ) > $O/r02b_pytest.log 2>&1
tail -12 $O/r02b_pytest.log
python bench.py --workload c3-raw --steps 10 --no-cpu > $O/r02b_bench_c3raw.json 2> $O/r02b_bench_c3raw.err; tail -c 400 $O/r02b_bench_c3raw.err
python bench.py --workload c3-raw --steps 10 --no-cpu --no-e2e --split-sync > $O/r02b_bench_c3raw_split.json 2> $O/r02b_split.err; tail -c 400 $O/r02b_split.err
GF3_XCORR_PATH=split python bench.py --workload c3-raw --steps 10 --no-cpu --no-e2e --split-sync > $O/r02b_bench_c3raw_twokernel.json 2> $O/r02b_twok.err; tail -c 400 $O/r02b_twok.err
python tools/bench_stages.py > $O/r02b_stages.txt 2>&1; tail -12 $O/r02b_stages.txt
ncu --set full --clock-control none --import-source on -k regex:"xcorr_fused" -s 2 -c 1 -o $O/r02b_prof_xcorr_fused python bench.py --workload c3-raw --streams 1024 --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/r02b_ncu.log 2>&1
tail -3 $O/r02b_ncu.log
python - <<'PY'
import json
for f in ["r02b_bench_c3raw.json","r02b_bench_c3raw_split.json","r02b_bench_c3raw_twokernel.json"]:
    try:
        d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
        print(f, "%.1f Gbit/s" % (d["value"]/1e3), d["roofline"]["stages_ms"], "frac", d["roofline"]["frac"], d.get("parity"), (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
