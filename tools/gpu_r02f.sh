#!/bin/bash
# Round 2, GPU call F: suite (short-window exact phases), raw-stream bench direct vs staged receive, matched filter with
# the overlapping half kept in registers.
O=gpurun_out
( time python -m pytest tests -q -m gpu -rA -p no:cacheprovider ) > $O/r02f_pytest.log 2>&1
tail -6 $O/r02f_pytest.log
python bench.py --workload c3-raw --steps 10 --no-cpu --no-e2e > $O/r02f_c3raw.json 2> $O/r02f_c3raw.err; tail -c 300 $O/r02f_c3raw.err
GF3_STREAMS_STAGED=1 python bench.py --workload c3-raw --steps 10 --no-cpu --no-e2e > $O/r02f_c3raw_staged.json 2> $O/r02f_c3raw_staged.err; tail -c 300 $O/r02f_c3raw_staged.err
GF3_RX_STAGED=1 python bench.py --steps 30 --no-cpu --no-e2e > $O/r02f_bench_c3_staged.json 2> $O/r02f_c3s.err; tail -c 300 $O/r02f_c3s.err
python bench.py --steps 30 --no-cpu > $O/r02f_bench_c3.json 2> $O/r02f_c3.err; tail -c 300 $O/r02f_c3.err
python - <<'PY'
import json
for f in ["r02f_c3raw.json","r02f_c3raw_staged.json","r02f_bench_c3_staged.json","r02f_bench_c3.json"]:
    try:
        d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
        r=d["roofline"]
        p=d.get("parity") or {}
        print(f, "%.1f Gbit/s" % (d["value"]/1e3), "ms/step %.3f" % d["ms_per_step"], "frac %.3f" % r["frac"], r.get("stages_ms"), "parity mism", p.get("bit_mismatches"), "beyond", p.get("beyond"), "e2e par", (p.get("e2e") or {}).get("beyond"),
              "e2e", [(k, round(d[k]["value"]/1e3,1), d[k].get("matches_device_result")) for k in ("e2e","e2e_pcm16","e2e_f32") if d.get(k) and d[k].get("value")])
    except Exception as e:
        print(f, "ERR", e)
PY
