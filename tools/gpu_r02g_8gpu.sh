#!/bin/bash
# Round 2, 8-GPU call: configs[4] sweep at 1/2/4/8 ranks (identical counters at every N) + the C3 bench at 8 ranks.
O=gpurun_out
nvidia-smi topo -m > $O/r02g_topo.txt 2>&1
for n in 1 2 4 8; do
  if [ $n = 1 ]; then
    python gf3-audio-modem_b200/gf3b200/sweep.py --streams 4096 > $O/r02g_sweep_n$n.json 2> $O/r02g_sweep_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) gf3-audio-modem_b200/gf3b200/sweep.py --streams 4096 > $O/r02g_sweep_n$n.json 2> $O/r02g_sweep_n$n.err
  fi
  tail -c 200 $O/r02g_sweep_n$n.err
  python -c "import json,sys; d=json.loads(open('$O/r02g_sweep_n$n.json').read().strip().splitlines()[-1]); print('sweep N=$n', d['seconds'], d['counters_sha256'][:16], d['sync_failures'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r02g_bench_c3_8gpu.json 2> $O/r02g_bench_8gpu.err
tail -c 300 $O/r02g_bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02g_bench_c3_8gpu.json").read().strip().splitlines()[-1])
print("8 GPU: %.1f Gbit/s, frac %.3f" % (d["value"]/1e3, d["roofline"]["frac"]), [(k, round(d[k]["value"]/1e3,1), round(d[k].get("h2d_gbs_per_gpu",0),1), round(d[k].get("h2d_ceiling_gbs_per_gpu",0),1)) for k in ("e2e","e2e_pcm16","e2e_f32") if d.get(k) and d[k].get("value")], d.get("numa"))
PY
