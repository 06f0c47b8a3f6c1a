#!/bin/bash
# Round 2, GPU call K: detection-only sync (bound-and-skip matched filter): parity + timing.
O=gpurun_out
( time python -m pytest tests/test_gpu_scale_parity.py tests/test_gpu_parity.py -q -m gpu -rA -p no:cacheprovider -k "sync or quirk or peak or roundtrip or sweep" ) > $O/r02k_pytest.log 2>&1
tail -8 $O/r02k_pytest.log
python bench.py --workload c3-raw --steps 10 --no-cpu --no-e2e > $O/r02k_c3raw.json 2> $O/r02k_c3raw.err; tail -c 300 $O/r02k_c3raw.err
python bench.py --workload c3-raw --steps 10 --no-cpu --no-e2e --dense-sync > $O/r02k_c3raw_dense.json 2> $O/r02k_c3raw_dense.err; tail -c 300 $O/r02k_c3raw_dense.err
python - <<'PY'
import json
for f in ["r02k_c3raw.json","r02k_c3raw_dense.json"]:
    try:
        d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
        r=d["roofline"]; p=d.get("parity") or {}
        print(f, "%.1f Gbit/s" % (d["value"]/1e3), "ms/step %.3f" % d["ms_per_step"], r.get("stages_ms"), "parity mism", p.get("bit_mismatches"), "sync mism", p.get("sync_index_mismatches"), "sync fail", d["check"]["streams_sync_failed"])
    except Exception as e:
        print(f, "ERR", e)
PY
