#!/bin/bash
# One GPU call of round 2: tests, A/B timings of the kernel variants (env toggles), bench line,
# ncu launch list + full captures.  Everything lands in gpurun_out/r02d_*.
mkdir -p gpurun_out
O=gpurun_out
B200NN_TRACE=1 timeout 240 python tools/multi_emul.py 2 1000000 > $O/r02d_multi_emul.log 2>&1; echo "multi emul rc=$?"; tail -6 $O/r02d_multi_emul.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02d_pytest.log; tail -4 $O/r02d_pytest.log
{
echo "== SNN stage 1M x 50 k=20: default (staged rows, symmetric, sorted suffix) | unstaged (filter2) | no symmetric | round-1 kernel, no symmetric"
python tools/prof_snn.py --reps 3
B200NN_SNN_UNSTAGED=1 python tools/prof_snn.py --reps 3
B200NN_SNN_NO_SYM=1 python tools/prof_snn.py --reps 3
B200NN_SNN_ROUND1=1 B200NN_SNN_NO_SYM=1 python tools/prof_snn.py --reps 3
echo "== SNN stage 1M x 30 k=30: default | unstaged"
python tools/prof_snn.py --n 1000000 --d 30 --k 30 --seed 2 --reps 2
B200NN_SNN_UNSTAGED=1 python tools/prof_snn.py --n 1000000 --d 30 --k 30 --seed 2 --reps 2
echo "== kNN stage 1M x 50: default | FMA group assignment | arrival order inside groups"
python tools/knn_loop.py 0 4
B200NN_KM_ASSIGN_FMA=1 python tools/knn_loop.py 0 4
} > $O/r02d_ab.txt 2>&1
tail -60 $O/r02d_ab.txt
timeout 420 python bench.py --steps 5 --warmup 3 > $O/r02d_bench_C.json 2> $O/r02d_bench_C.err; echo "bench rc=$?"; tail -c 1500 $O/r02d_bench_C.json
timeout 300 python bench.py --impl reference > $O/r02d_bench_ref.json 2> $O/r02d_bench_ref.err; echo "ref rc=$?"
# ncu: launch list of one short bench run, then full captures of the two top kernels
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-verify --no-extra-workload --e2e-steps 1 > $O/r02d_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r02d_launches_bench_C.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-verify --no-extra-workload --e2e-steps 1 > $O/r02d_ncu_bench.log 2>&1
python tools/prof_snn.py --reps 2 > $O/r02d_plain_snn.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_snn_rows -s 3 -c 3 -o $O/r02d_k_snn_rows python tools/prof_snn.py --reps 2 > $O/r02d_ncu_snn.log 2>&1
ls -la $O | grep r02d; true
