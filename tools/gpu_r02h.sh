#!/bin/bash
# Round 2, GPU call H: full GPU suite (legacy API shims, device-side XOR encode) + smoke.
O=gpurun_out
( time python -m pytest tests -q -m gpu -rA -p no:cacheprovider ) > $O/r02h_pytest.log 2>&1
tail -8 $O/r02h_pytest.log
python __graft_entry__.py smoke > $O/r02h_smoke.log 2>&1; tail -2 $O/r02h_smoke.log
python tools/time_kat1.py > $O/r02h_kat1.log 2>&1; tail -3 $O/r02h_kat1.log
