#!/bin/bash
# Round-2 third session, 1-GPU check of the committed state: the whole GPU suite with durations,
# the smoke call, the default bench line.  Every step has its own time-out; logs: gpurun_out/r02g_*.
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader > $O/r02g_box.txt 2>&1; nproc >> $O/r02g_box.txt
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q --durations=20 -o faulthandler_timeout=600 > $O/r02g_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02g_pytest.log; tail -30 $O/r02g_pytest.log
timeout 120 python __graft_entry__.py --smoke > $O/r02g_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02g_smoke.log
timeout 420 python bench.py --steps 5 --warmup 3 > $O/r02g_bench_C.json 2> $O/r02g_bench_C.err; echo "bench rc=$?"; tail -c 3000 $O/r02g_bench_C.json; tail -5 $O/r02g_bench_C.err
