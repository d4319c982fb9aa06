#!/bin/bash
# final multi-GPU check with the committed code: the bench line as the driver launches it (value + e2e)
N=${1:-2}
O=gpurun_out
mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 \
    bench.py --gpus $N --steps 5 --warmup 3 > $O/r02q_bench_C_${N}gpu.json 2> $O/r02q_bench_C_${N}gpu.err
echo "bench C $N gpus rc=$?"; grep "^{" $O/r02q_bench_C_${N}gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'verify ok', d['verify']['ok'], d['stage_ms_per_step'])
e=d['e2e']; print('e2e', {k:v for k,v in e.items() if k!='api'})"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --impl reference --gpus $N --steps 5 --warmup 3 > $O/r02q_bench_ref_${N}gpu.json 2> $O/r02q_bench_ref_${N}gpu.err
echo "reference arm rc=$?"; grep "^{" $O/r02q_bench_ref_${N}gpu.json | head -c 400; echo
