#!/bin/bash
# edge-file input of the modularity optimiser, resident index across split counts, fallback slices
O=gpurun_out
mkdir -p $O
timeout 300 python -X faulthandler -m pytest tests/test_gpu_modopt.py tests/test_gpu_large.py -x -q --durations=6 -k "edge_file or resident_index or known_answers" > $O/r02k_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02k_pytest.log; tail -12 $O/r02k_pytest.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra-workload --no-clustering --e2e-steps 2 > $O/r02k_bench.json 2> $O/r02k_bench.err
echo "bench rc=$?"; python - <<PY
import json
d=json.load(open("$O/r02k_bench.json"))
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "uncertified", d["n_uncertified"], d["stage_ms_per_step"], "verify", d["verify"]["ok"])
PY
