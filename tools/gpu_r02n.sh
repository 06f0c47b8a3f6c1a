#!/bin/bash
O=gpurun_out
( time python -m pytest tests -q -m gpu -rA -p no:cacheprovider ) > $O/r02n_pytest.log 2>&1
tail -6 $O/r02n_pytest.log
