#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_scale_parity.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "long_chirp or a2 or kat1 or kat4 or sync or xcorr or uint8 or pcm" 2>&1 | tail -3
timeout 300 python bench.py --workload a2-raw --steps 20 --no-cpu > $O/r02ao_bench_a2raw.json 2> $O/r02ao.err || tail -c 600 $O/r02ao.err
python -c "import json; d=json.loads(open('$O/r02ao_bench_a2raw.json').read().strip().splitlines()[-1]); r=d.get('roofline') or {}; print(round(d['value']/1e3,2),'Gbit/s', round(d['ms_per_step'],3),'ms', r.get('stages_ms'), (d.get('parity') or {}).get('bit_mismatches'), [round(d[k]['value']/1e3,1) for k in ('e2e','e2e_pcm16','e2e_f32')])"
