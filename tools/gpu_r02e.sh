#!/bin/bash
# Round 2, GPU call E: full GPU suite after the narrow-load fix; ncu captures (fused matched filter at 2 CTAs/SM,
# staged float32 C3 kernel for the TMA question); raw-stream bench with staged receive.
O=gpurun_out
python tools/debug_staged.py > $O/r02e_debug_staged.log 2>&1; grep -c "rc 0" $O/r02e_debug_staged.log; grep "rc 1" $O/r02e_debug_staged.log | head
( time python -m pytest tests -q -m gpu -rA -p no:cacheprovider ) > $O/r02e_pytest.log 2>&1
tail -8 $O/r02e_pytest.log
python bench.py --workload c3-raw --steps 10 --no-cpu > $O/r02e_c3raw.json 2> $O/r02e_c3raw.err; tail -c 300 $O/r02e_c3raw.err
GF3_STREAMS_STAGED=1 python bench.py --workload c3-raw --steps 10 --no-cpu --no-e2e > $O/r02e_c3raw_staged.json 2> $O/r02e_c3raw_staged.err; tail -c 300 $O/r02e_c3raw_staged.err
ncu --set full --clock-control none --import-source on -k regex:"xcorr_fused" -s 2 -c 1 -o $O/r02e_prof_xcorr_fused python bench.py --workload c3-raw --streams 1024 --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/r02e_ncu1.log 2>&1
GF3_RX_STAGED=1 ncu --set full --clock-control none --import-source on -k regex:"rx_demod" -s 3 -c 1 -o $O/r02e_prof_c3_staged python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/r02e_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rx_demod" -s 3 -c 1 -o $O/r02e_prof_c3 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/r02e_ncu3.log 2>&1
python - <<'PY'
import json
for f in ["r02e_c3raw.json","r02e_c3raw_staged.json"]:
    try:
        d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, "%.1f Gbit/s" % (d["value"]/1e3), "ms/step %.3f" % d["ms_per_step"], "frac %.3f" % r["frac"], r.get("stages_ms"), "parity", d.get("parity"),
              "e2e", [(k, round(d[k]["value"]/1e3,1), d[k].get("matches_device_result")) for k in ("e2e","e2e_pcm16","e2e_f32") if d.get(k) and d[k].get("value")])
    except Exception as e:
        print(f, "ERR", e)
PY
