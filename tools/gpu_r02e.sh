#!/bin/bash
# Round-2 re-entry check on one GPU: emulated multi-shard run with the phase trace, the whole GPU
# suite with per-test durations, the default bench line and the reference arm.  Every step has its
# own time-out; logs land in gpurun_out/r02e_*.
O=gpurun_out
mkdir -p $O
B200NN_TRACE=1 timeout 200 python tools/multi_emul.py 2 300000 > $O/r02e_multi_emul.log 2>&1; echo "multi emul rc=$?"; tail -4 $O/r02e_multi_emul.log
timeout 1500 python -X faulthandler -m pytest tests -m gpu -x -q --durations=25 -o faulthandler_timeout=600 > $O/r02e_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02e_pytest.log; tail -40 $O/r02e_pytest.log
timeout 420 python bench.py --steps 5 --warmup 3 > $O/r02e_bench_C.json 2> $O/r02e_bench_C.err; echo "bench rc=$?"; tail -c 2500 $O/r02e_bench_C.json; tail -5 $O/r02e_bench_C.err
timeout 300 python bench.py --impl reference > $O/r02e_bench_ref.json 2> $O/r02e_bench_ref.err; echo "ref rc=$?"; tail -c 600 $O/r02e_bench_ref.json
timeout 120 python __graft_entry__.py --smoke > $O/r02e_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02e_smoke.log
