#!/bin/bash
# epilogue chunk-buffer depth sweep (TFMQ_IGEMM_EPI_BUFS): conv microbench + bench line per depth
mkdir -p gpurun_out
tag=${1:-eb}
for nb in 0 3 4 5 6; do
  TFMQ_IGEMM_EPI_BUFS=$nb timeout 300 python tools/microbench_conv.py 2>&1 | grep -v "i8 gemm" > gpurun_out/${tag}_micro_nb$nb.txt
done
for nb in 0 4 6 0; do
  TFMQ_IGEMM_EPI_BUFS=$nb timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('nb=$nb ms/step %.3f e2e %.3f w4a8 %.3f ms frac %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac']))"
done
python - <<PY
import re
cols = {}
for nb in (0, 3, 4, 5, 6):
    for l in open("gpurun_out/${tag}_micro_nb%d.txt" % nb):
        m = re.match(r"(.*):\s+([0-9.]+) us", l)
        if m: cols.setdefault(m.group(1), []).append(m.group(2))
for k, v in cols.items(): print("%-62s" % k, " ".join("%7s" % x for x in v))
PY
