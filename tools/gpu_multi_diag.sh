#!/bin/bash
# short diagnostic of the sharded-path tests on a multi-GPU box: per-test time-outs with Python
# tracebacks (faulthandler) and the phase trace of csrc/multi.cu
N=${1:-2}
O=gpurun_out
mkdir -p $O
export B200NN_TRACE=1
for t in test_multi_matches_single_device test_multi_through_the_reference_interface test_multi_edge_cases \
         test_multi_config_c_all_devices test_dev_knn_shares_partition_the_rows test_nccl_ranks_match_single_gpu; do
  timeout 150 python -X faulthandler -m pytest tests/test_gpu_multi.py -x -q -k $t -o faulthandler_timeout=100 \
      > $O/r02_diag_${t}.log 2>&1
  echo "$t rc=$?"; grep -E "passed|failed|error|Timeout|rc=" $O/r02_diag_${t}.log | tail -3
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 5 --warmup 3 > $O/r02_bench_C_${N}gpu.json 2> $O/r02_bench_C_${N}gpu.err
echo "bench $N gpus rc=$?"; tail -c 800 $O/r02_bench_C_${N}gpu.json
