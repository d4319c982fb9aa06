#!/bin/bash
# final check of round 2 on one GPU: whole GPU suite, smoke, the default bench line, the reference arm,
# then (each only after its own command exited 0 without ncu) the ncu launch list of the bench command and a
# full capture of the three k_knn_tc launches of one kNN call (group assignment, pass 1, pass 2)
O=gpurun_out
mkdir -p $O
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q --durations=12 -o faulthandler_timeout=600 > $O/r02p_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02p_pytest.log; tail -18 $O/r02p_pytest.log
timeout 120 python __graft_entry__.py --smoke > $O/r02p_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02p_smoke.log
timeout 420 python bench.py --steps 5 --warmup 3 > $O/r02p_bench_C.json 2> $O/r02p_bench_C.err; echo "bench rc=$?"; head -c 900 $O/r02p_bench_C.json; echo
SHORT="--steps 2 --warmup 1 --no-cpu-baseline --no-verify --no-extra-workload --no-clustering --e2e-steps 1"
timeout 200 python bench.py $SHORT > $O/r02p_plain_bench.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r02p_launches_bench_C.csv \
    python bench.py $SHORT > $O/r02p_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 100 python tools/knn_loop.py 0 4 > $O/r02p_plain_knn.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_knn_tc --launch-skip 24 --launch-count 3 -o $O/r02p_k_knn_tc \
    python tools/knn_loop.py 0 4 > $O/r02p_ncu_knn.log 2>&1; echo "ncu full rc=$?"; tail -3 $O/r02p_plain_knn.log
ls -la $O | grep r02p
