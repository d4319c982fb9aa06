#!/bin/bash
# Second GPU call of the re-entry session: first run of the modularity optimiser on hardware, the
# multi-device suite after the non-finite fix, SNN staged / unstaged A/B.  Logs: gpurun_out/r02f_*.
O=gpurun_out
mkdir -p $O
timeout 300 python -X faulthandler -m pytest tests/test_gpu_modopt.py -x -q --durations=10 > $O/r02f_pytest_modopt.log 2>&1; echo "pytest modopt rc=$?"; tail -15 $O/r02f_pytest_modopt.log
timeout 150 python tools/prof_modopt.py --n 1000000 --starts 1 --iters 10 --reps 2 --ref > $O/r02f_modopt_1m.log 2>&1; echo "modopt 1M rc=$?"; tail -5 $O/r02f_modopt_1m.log
B200NN_TRACE=1 timeout 400 python -X faulthandler -m pytest tests/test_gpu_multi.py -x -q --durations=10 > $O/r02f_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; grep -v "b200nn multi" $O/r02f_pytest_multi.log | tail -15
{
echo "== SNN stage 1M x 50 k=20: default (unstaged filter2) | staged"
timeout 60 python tools/prof_snn.py --reps 3
B200NN_SNN_STAGED=1 timeout 60 python tools/prof_snn.py --reps 3
} > $O/r02f_snn_ab.txt 2>&1
tail -12 $O/r02f_snn_ab.txt
