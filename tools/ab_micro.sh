#!/bin/bash
# A/B of lib/libA.so vs lib/libB.so on the conv microbench (graph-replayed GPU times)
L=tfmq-dm_b200/tfmq_b200/lib
tag=${1:-abm}
mkdir -p gpurun_out
for v in A B; do
  cp $L/lib$v.so $L/libtfmq_b200.so
  TFMQ_ONLY_GEMM= timeout 300 python tools/microbench_conv.py 2>&1 | grep -v "i8 gemm" > gpurun_out/${tag}_micro_$v.txt
  timeout 300 python tools/microbench_small_maps.py 2>&1 | grep -v TFMQ >> gpurun_out/${tag}_micro_$v.txt
done
python - <<PY
import re
cols = {}
for v in "AB":
    for l in open("gpurun_out/${tag}_micro_%s.txt" % v):
        m = re.match(r"(.*):\s+([0-9.]+) us", l)
        if m: cols.setdefault(m.group(1), []).append(m.group(2))
        elif l.strip(): print(v, l.strip()[:300])
for k, v in cols.items(): print("%-62s" % k, " ".join("%7s" % x for x in v))
PY
