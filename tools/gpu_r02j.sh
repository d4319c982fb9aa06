#!/bin/bash
# column splits with binomial head room: the k > 32 tests and the large-k timing at 1M
O=gpurun_out
mkdir -p $O
timeout 400 python -X faulthandler -m pytest tests/test_gpu_tc.py -x -q --durations=8 -s -k "large_k or k50" > $O/r02j_pytest_tc.log 2>&1; echo "pytest tc rc=$?" | tee -a $O/r02j_pytest_tc.log; grep -E "uncertified|passed|failed|Error|error" $O/r02j_pytest_tc.log | tail -30
timeout 300 python tools/prof_large_k.py --ks 33,50,100,200 > $O/r02j_large_k.txt 2>&1; echo "large k rc=$?"; cat $O/r02j_large_k.txt | tail -6
