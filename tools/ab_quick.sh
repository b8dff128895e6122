#!/bin/bash
# quick A/B on one box: kernel tests with lib/libB.so, then one bench line each for lib/libA.so and lib/libB.so
L=tfmq-dm_b200/tfmq_b200/lib
cp $L/libB.so $L/libtfmq_b200.so
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -n 2
for v in A B; do
  cp $L/lib$v.so $L/libtfmq_b200.so
  timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra 2>/dev/null | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v ms/step %.3f e2e %.3f w4a8 %.3f ms frac %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac']))"
done
cp $L/libB.so $L/libtfmq_b200.so
