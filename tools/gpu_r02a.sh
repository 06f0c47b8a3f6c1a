#!/bin/bash
# Round 2, GPU call A: full GPU test suite, bench lines of every workload, sweep on 1 GPU, sanitizer runs.
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02a_smi.csv 2>&1
( time python -m pytest tests -q -m gpu -rA -p no:cacheprovider ) > $O/r02a_pytest.log 2>&1
tail -15 $O/r02a_pytest.log
python bench.py > $O/r02a_bench_c3.json 2> $O/r02a_bench_c3.err; tail -c 600 $O/r02a_bench_c3.err
python bench.py --workload c3-raw --steps 10 > $O/r02a_bench_c3raw.json 2> $O/r02a_bench_c3raw.err; tail -c 600 $O/r02a_bench_c3raw.err
python bench.py --workload c4 --no-cpu --steps 30 > $O/r02a_bench_c4.json 2> $O/r02a_bench_c4.err; tail -c 300 $O/r02a_bench_c4.err
python bench.py --workload c4-long --no-cpu --steps 30 > $O/r02a_bench_c4long.json 2> $O/r02a_bench_c4long.err; tail -c 300 $O/r02a_bench_c4long.err
python bench.py --workload a2 --no-cpu --steps 30 > $O/r02a_bench_a2.json 2> $O/r02a_bench_a2.err; tail -c 300 $O/r02a_bench_a2.err
python bench.py --impl reference --workload c3-raw --steps 3 --warmup 1 > $O/r02a_bench_ref_c3raw.json 2> $O/r02a_bench_ref.err
( time python gf3-audio-modem_b200/gf3b200/sweep.py --streams 1024 ) > $O/r02a_sweep_n1.json 2> $O/r02a_sweep_n1.err; tail -c 300 $O/r02a_sweep_n1.err
python tools/bench_stages.py > $O/r02a_stages.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py > $O/r02a_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r02a_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_smoke.py > $O/r02a_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/r02a_racecheck.log
tail -3 $O/r02a_memcheck.log $O/r02a_racecheck.log
head -c 1500 $O/r02a_bench_c3.json; echo; head -c 1200 $O/r02a_bench_c3raw.json
