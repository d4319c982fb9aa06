#!/bin/bash
# multi-GPU record of round 2 on N devices: the bench line as the driver launches it (value + multi-device e2e
# with the phase trace), configs D and E sharded through the single-process multi-device C ABI
N=${1:-4}
O=gpurun_out
mkdir -p $O
B200NN_TRACE=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
    bench.py --gpus $N --steps 5 --warmup 3 > $O/r02m_bench_C_${N}gpu.json 2> $O/r02m_bench_C_${N}gpu.err
echo "bench C $N gpus rc=$?"; grep "^{" $O/r02m_bench_C_${N}gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'verify', d['verify'], d['stage_ms_per_step'])
e=d['e2e']; print('e2e', {k:v for k,v in e.items() if k!='api'})"
grep "b200nn multi" $O/r02m_bench_C_${N}gpu.err | tail -$((N*14)) | grep "shard 0/" | tail -14
timeout 300 python tools/run_configs.py D --multi $N > $O/r02m_config_D_multi_${N}gpu.json 2> $O/r02m_config_D_multi_${N}gpu.err; echo "D multi rc=$?"; head -c 1500 $O/r02m_config_D_multi_${N}gpu.json; echo
timeout 240 python tools/run_configs.py E --multi $N > $O/r02m_config_E_multi_${N}gpu.json 2> $O/r02m_config_E_multi_${N}gpu.err; echo "E multi rc=$?"; head -c 1200 $O/r02m_config_E_multi_${N}gpu.json; echo
