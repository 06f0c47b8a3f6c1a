#!/bin/bash
O=gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -4
python tools/bench_stages.py > $O/r02y_stages.md 2>&1; tail -11 $O/r02y_stages.md
