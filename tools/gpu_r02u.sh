#!/bin/bash
# N = 4096: last FFT pass on adjacent columns (GF3_RX12_ALT=3) against the default build, same box
O=gpurun_out
line() { python -c "import json; d=json.loads(open('$1').read().strip().splitlines()[-1]); r=d['roofline']; p=d['parity']; print('$2', round(d['value']/1e3,1), 'Gbit/s ms/step', round(d['ms_per_step'],4), 'kernel', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],4), 'chain', round(r['chain']['frac'],4), 'parity', p['bit_mismatches'], p['beyond'], p.get('within_eq_tol'))"; }
for v in b200 _alt12p b200 _alt12p; do
  export GF3_LIB_PATH=$PWD/gf3-audio-modem_b200/lib/libgf3$v.so
  for w in c4 a2 c4-long; do
    python bench.py --workload $w --steps 30 --no-cpu --no-e2e > $O/r02u_${w}$v.json 2> $O/r02u$v.err
    line $O/r02u_${w}$v.json "$v $w" || tail -c 300 $O/r02u$v.err
  done
done
export GF3_LIB_PATH=$PWD/gf3-audio-modem_b200/lib/libgf3_alt12p.so
python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:"rx_demod" -s 3 -c 1 -o $O/r02u_prof_c4_alt12p python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/r02u_ncu.log 2>&1
tail -1 $O/r02u_ncu.log
