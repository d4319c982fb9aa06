#!/bin/bash
# async gather in the grouped FP64 fallback: uncertified-row tests + bench stage times
O=gpurun_out
mkdir -p $O
timeout 300 python -X faulthandler -m pytest tests/test_gpu_tc.py -x -q -k "uncertified or awkward or ties or badly" > $O/r02l_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02l_pytest.log; tail -3 $O/r02l_pytest.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra-workload --no-clustering --e2e-steps 2 > $O/r02l_bench.json 2> $O/r02l_bench.err
echo "bench rc=$?"; python - <<PY
import json
d=json.load(open("$O/r02l_bench.json"))
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "uncertified", d["n_uncertified"], d["stage_ms_per_step"], "verify", d["verify"]["ok"])
PY
