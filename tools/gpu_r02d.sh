#!/bin/bash
# Round 2, GPU call D: staged-input (cp.async.bulk) receive kernels -- parity (PCM + float32) and timing against
# the direct-load path; fused matched-filter variants; full GPU suite.
O=gpurun_out
( time python -m pytest tests -q -m gpu -rA -p no:cacheprovider ) > $O/r02d_pytest.log 2>&1
tail -12 $O/r02d_pytest.log
python bench.py --steps 30 --no-cpu > $O/r02d_bench_c3.json 2> $O/r02d_c3.err; tail -c 300 $O/r02d_c3.err
GF3_RX_STAGED=1 python bench.py --steps 30 --no-cpu --no-e2e > $O/r02d_bench_c3_staged.json 2> $O/r02d_c3s.err; tail -c 300 $O/r02d_c3s.err
python bench.py --workload c4 --steps 30 --no-cpu --no-e2e > $O/r02d_bench_c4.json 2> $O/r02d_c4.err; tail -c 300 $O/r02d_c4.err
GF3_RX_STAGED=1 python bench.py --workload c4 --steps 30 --no-cpu --no-e2e > $O/r02d_bench_c4_staged.json 2> $O/r02d_c4s.err; tail -c 300 $O/r02d_c4s.err
for mb in 2 3; do
  GF3_XC_MINB=$mb python bench.py --workload c3-raw --steps 10 --no-cpu --no-e2e --no-parity --split-sync > $O/r02d_c3raw_minb$mb.json 2> $O/r02d_minb$mb.err; tail -c 300 $O/r02d_minb$mb.err
done
GF3_STREAMS_STAGED=1 python bench.py --workload c3-raw --steps 10 --no-cpu > $O/r02d_c3raw_staged.json 2> $O/r02d_c3raw_staged.err; tail -c 300 $O/r02d_c3raw_staged.err
python - <<'PY'
import json
for f in ["r02d_bench_c3.json","r02d_bench_c3_staged.json","r02d_bench_c4.json","r02d_bench_c4_staged.json","r02d_c3raw_minb2.json","r02d_c3raw_minb3.json","r02d_c3raw_staged.json"]:
    try:
        d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, "%.1f Gbit/s" % (d["value"]/1e3), "ms/step %.3f" % d["ms_per_step"], "frac %.3f" % r["frac"], r.get("stages_ms"), "parity", (d.get("parity") or {}).get("bit_mismatches"), (d.get("parity") or {}).get("beyond"),
              "e2e", [(k, round(d[k]["value"]/1e3,1), d[k].get("matches_device_result")) for k in ("e2e","e2e_pcm16","e2e_f32") if d.get(k) and d[k].get("value")])
    except Exception as e:
        print(f, "ERR", e)
PY
