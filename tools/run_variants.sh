for v in t256p0 t128p0 t256p1 t128p2 t256p2 t64p0; do
  L=$PWD/gf3-audio-modem_b200/lib/libgf3b200_$v.so
  GF3_LIB_PATH=$L python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stage_receive or loopback" 2>&1 | tail -1
  GF3_LIB_PATH=$L python bench.py --no-cpu --no-e2e --steps 30 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('variant','$v','ms/step',round(d['ms_per_step'],4),'demod ms',round(r['avg_launch_ms'],4),'frac',round(r['frac'],4))"
done
L=$PWD/gf3-audio-modem_b200/lib/libgf3b200_t128p0.so
GF3_LIB_PATH=$L ncu --set full --clock-control none --import-source on -k regex:rx_demod -s 3 -c 1 -o gpurun_out/prof_demod_t128p0 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log
