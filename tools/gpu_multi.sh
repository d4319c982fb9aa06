#!/bin/bash
# multi-GPU call: sharded-path tests, the bench line at N = 2, 4, 8 (as the driver launches it) and
# configs D / E through the single-process multi-device C ABI.  Usage: gpu_multi.sh <n_gpus>
# Every step runs under its own `timeout` so that one hanging step cannot hold the box.
N=${1:-8}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r02_topo_${N}gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > $O/r02_pytest_multi_${N}gpu.log 2>&1; echo "pytest multi rc=$?" >> $O/r02_pytest_multi_${N}gpu.log; tail -5 $O/r02_pytest_multi_${N}gpu.log
for G in 2 4 8; do
  if [ $G -le $N ]; then
    timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2951$G \
        bench.py --gpus $G --steps 5 --warmup 3 > $O/r02_bench_C_${G}gpu.json 2> $O/r02_bench_C_${G}gpu.err
    echo "bench $G gpus rc=$?"; tail -c 600 $O/r02_bench_C_${G}gpu.json
  fi
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 \
    bench.py --gpus $N --config D --steps 3 --warmup 2 > $O/r02_bench_D_${N}gpu.json 2> $O/r02_bench_D_${N}gpu.err; echo "bench D rc=$?"
timeout 600 python tools/run_configs.py D --multi $N > $O/r02_config_D_multi_${N}gpu.json 2> $O/r02_config_D_multi_${N}gpu.err; echo "D multi rc=$?"; head -c 800 $O/r02_config_D_multi_${N}gpu.json
timeout 420 python tools/run_configs.py E --multi $N > $O/r02_config_E_multi_${N}gpu.json 2> $O/r02_config_E_multi_${N}gpu.err; echo "E multi rc=$?"; head -c 800 $O/r02_config_E_multi_${N}gpu.json
