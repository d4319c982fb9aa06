#!/bin/bash
# column splits (k > 32): tensor-path suite, large-k timing at 1M, candidate-list width A/B on the bench line
O=gpurun_out
mkdir -p $O
timeout 400 python -X faulthandler -m pytest tests/test_gpu_tc.py -x -q --durations=8 -s > $O/r02i_pytest_tc.log 2>&1; echo "pytest tc rc=$?" | tee -a $O/r02i_pytest_tc.log; grep -E "uncertified|passed|failed|Error|error" $O/r02i_pytest_tc.log | tail -30
timeout 200 python tools/prof_large_k.py > $O/r02i_large_k.txt 2>&1; echo "large k rc=$?"; cat $O/r02i_large_k.txt | tail -5
for NC in 24 28; do
  timeout 200 python bench.py --steps 5 --warmup 3 --n-cand $NC --no-cpu-baseline --no-extra-workload --no-clustering --e2e-steps 1 > $O/r02i_bench_ncand$NC.json 2> $O/r02i_bench_ncand$NC.err
  echo "bench n_cand=$NC rc=$?"; python - <<PY
import json
d=json.load(open("$O/r02i_bench_ncand$NC.json"))
print("ms_per_step", d["ms_per_step"], "uncertified", d["n_uncertified"], d["stage_ms_per_step"], "verify", d["verify"]["ok"])
PY
done
