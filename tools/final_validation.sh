#!/bin/bash
# round-end validation on one B200: GPU tests, smoke, default bench line, reference arm, launch list of one step
mkdir -p gpurun_out
tag=${1:-final}
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/${tag}_gputests.txt 2>&1
tail -n 6 gpurun_out/${tag}_gputests.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 2
( time timeout 900 python bench.py ) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -n 1 gpurun_out/${tag}_bench.json | cut -c1-1500; tail -n 4 gpurun_out/${tag}_bench.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
tail -n 1 gpurun_out/${tag}_bench_ref.json | cut -c1-600; tail -n 4 gpurun_out/${tag}_bench_ref.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py > /dev/null 2>&1
wc -l gpurun_out/${tag}_launches.csv
