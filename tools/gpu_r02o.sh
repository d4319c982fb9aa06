#!/bin/bash
# anchor group with enough tiles: share balance, tensor-path suite, bench line
O=gpurun_out
mkdir -p $O
B200NN_DEBUG=1 timeout 120 python tools/share_balance.py 4 > $O/r02o_share_balance.txt 2>&1; grep -E "^world" $O/r02o_share_balance.txt; grep "pass 2" $O/r02o_share_balance.txt | sed -n "1p;3p"
timeout 300 python -X faulthandler -m pytest tests/test_gpu_tc.py tests/test_gpu_large.py -x -q > $O/r02o_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02o_pytest.log; tail -3 $O/r02o_pytest.log
B200NN_DEBUG=1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-clustering --e2e-steps 2 > $O/r02o_bench.json 2> $O/r02o_bench.err
echo "bench rc=$?"; grep "pass 2" $O/r02o_bench.err | tail -2; python - <<PY
import json
d=json.load(open("$O/r02o_bench.json"))
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "uncertified", d["n_uncertified"], d["stage_ms_per_step"], "verify", d["verify"]["ok"])
print(d["roofline"]["frac"], d["roofline"]["keys_per_clk_per_sm"], d["other_workloads"])
PY
