#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
python bench.py > $O/r02at_bench_c3.json 2> $O/r02at.err; python -c "import json; d=json.loads(open('$O/r02at_bench_c3.json').read().strip().splitlines()[-1]); print(round(d['value']/1e3,1), d['roofline']['frac'], d['e2e']['value'], d['parity']['bit_mismatches'], d['cpu_baseline']['value'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
