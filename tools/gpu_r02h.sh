#!/bin/bash
# 2-GPU check: sharded-path tests on real peer devices, then the bench line as the driver launches it.
N=${1:-2}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r02h_topo_${N}gpu.txt 2>&1
B200NN_TRACE=1 timeout 400 python -X faulthandler -m pytest tests/test_gpu_multi.py -x -q --durations=10 -o faulthandler_timeout=300 > $O/r02h_pytest_multi_${N}gpu.log 2>&1; echo "pytest multi rc=$?" | tee -a $O/r02h_pytest_multi_${N}gpu.log; grep -v "b200nn multi" $O/r02h_pytest_multi_${N}gpu.log | tail -15
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 5 --warmup 3 > $O/r02h_bench_C_${N}gpu.json 2> $O/r02h_bench_C_${N}gpu.err
echo "bench $N gpus rc=$?"; tail -c 2500 $O/r02h_bench_C_${N}gpu.json; tail -5 $O/r02h_bench_C_${N}gpu.err
