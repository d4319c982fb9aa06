#!/bin/bash
# A/B timings of the round-2 kernels on config C (1M x 50): env toggles select variants.
mkdir -p gpurun_out
{
echo "== SNN stage: byte-flag single walk (8k flags) | 4k flags | no read-only lists | two-walk (round 1)"
python tools/prof_snn.py --reps 3
B200NN_SNN_LV1=12 python tools/prof_snn.py --reps 3
B200NN_SNN_NO_RO=1 python tools/prof_snn.py --reps 3
B200NN_SNN_TWO_WALK=1 python tools/prof_snn.py --reps 3
echo "== SNN stage, 4M x 30 k=30 (n 1000000 subset): default | two-walk"
python tools/prof_snn.py --n 1000000 --d 30 --k 30 --seed 2 --reps 2
B200NN_SNN_TWO_WALK=1 python tools/prof_snn.py --n 1000000 --d 30 --k 30 --seed 2 --reps 2
} > gpurun_out/r02_ab2.txt 2>&1
tail -40 gpurun_out/r02_ab2.txt
