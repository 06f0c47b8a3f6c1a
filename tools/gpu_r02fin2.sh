#!/bin/bash
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02fin_bench_c3_2gpu.json 2> $O/r02fin2.err || tail -c 800 $O/r02fin2.err
python -c "import json; d=json.loads(open('$O/r02fin_bench_c3_2gpu.json').read().strip().splitlines()[-1]); print('2gpu', round(d['value']/1e3,1),'Gbit/s', d['n_gpus'], round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']/1e3,1), d['parity']['bit_mismatches'], d['clocks'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/r02fin_bench_ref_2gpu.json 2>> $O/r02fin2.err; tail -c 400 $O/r02fin_bench_ref_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 -m gf3b200.sweep --streams 1024 2>&1 | tail -2 | cut -c1-600
