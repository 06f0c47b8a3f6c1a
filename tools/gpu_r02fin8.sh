#!/bin/bash
O=gpurun_out
export PYTHONPATH=$PWD/gf3-audio-modem_b200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r02fin_bench_c3_8gpu.json 2> $O/r02fin8.err || tail -c 800 $O/r02fin8.err
python -c "import json; d=json.loads(open('$O/r02fin_bench_c3_8gpu.json').read().strip().splitlines()[-1]); print('8gpu', round(d['value']/1e3,1),'Gbit/s', d['n_gpus'], round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']/1e3,1), d['parity']['bit_mismatches'], d['clocks'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload a2-raw --steps 20 --warmup 5 --no-cpu > $O/r02fin_bench_a2raw_8gpu.json 2>> $O/r02fin8.err
python -c "import json; d=json.loads(open('$O/r02fin_bench_a2raw_8gpu.json').read().strip().splitlines()[-1]); print('8gpu a2-raw', round(d['value']/1e3,1),'Gbit/s', d['n_gpus'], round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']/1e3,1), d['parity']['bit_mismatches'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 -m gf3b200.sweep --streams 4096 > $O/r02fin_sweep_n8.json 2>> $O/r02fin8.err; tail -c 300 $O/r02fin_sweep_n8.json
