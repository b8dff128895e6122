#!/bin/bash
# A/B on one box: lib/libA.so against lib/libB.so (tools/build_variant.sh), alternating bench runs + the conv microbench
L=tfmq-dm_b200/tfmq_b200/lib
mkdir -p gpurun_out
tag=${1:-ab}
cp $L/libB.so $L/libtfmq_b200.so
timeout 400 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/${tag}_testsB.txt
for v in A B A B; do
  cp $L/lib$v.so $L/libtfmq_b200.so
  timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 >> gpurun_out/${tag}_bench_$v.json
done
for v in A B; do
  cp $L/lib$v.so $L/libtfmq_b200.so
  timeout 300 python tools/microbench_conv.py > gpurun_out/${tag}_micro_$v.txt 2>&1
done
cat gpurun_out/${tag}_testsB.txt
python - <<PY
import json
for v in "AB":
    for l in open("gpurun_out/${tag}_bench_%s.json" % v):
        try: d = json.loads(l)
        except Exception: print(v, "bad line", l[:200]); continue
        print(v, "ms/step %.3f  e2e %.3f  w4a8 %.3f ms frac %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"]))
PY
paste -d'\n' gpurun_out/${tag}_micro_A.txt gpurun_out/${tag}_micro_B.txt
