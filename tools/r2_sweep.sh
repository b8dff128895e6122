#!/bin/bash
# round-2 opening measurements: baseline bench, the ring-depth variants left unmeasured in round 1, int8 peak protocol
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_base.json 2> gpurun_out/r2a_base.err
for cfg in "8 8" "6 6" "8 4" "4 8" "6 4"; do
  set -- $cfg
  TFMQ_IGEMM_USTAGES=$1 TFMQ_IGEMM_PSTAGES=$2 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_u$1p$2.json 2> gpurun_out/r2a_u$1p$2.err
done
python tools/int8_peak.py > gpurun_out/r2a_int8_peak.json 2> gpurun_out/r2a_int8_peak.err
tail -n 3 gpurun_out/r2a_*.json
