#!/bin/bash
# Everything the round needs from one GPU box: parity tests, smoke, both bench arms, ncu launch list + full capture.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 300 --timeout-method thread 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/plain_a.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu_a.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/plain_b.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_score_tc -s 6 -c 1 -o gpurun_out/prof_score \
      python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/ncu_b.log 2>&1
python tools/kernel_time.py 100 100000 128 lutm > gpurun_out/plain_c.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_score_tc -s 3 -c 1 -o gpurun_out/prof_lutm \
      python tools/kernel_time.py 100 100000 128 lutm > gpurun_out/ncu_c.log 2>&1
python tools/closure_time.py > gpurun_out/closure_time.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -1; cat gpurun_out/bench.json | cut -c1-1500; cat gpurun_out/bench_reference.json | cut -c1-600; tail -2 gpurun_out/bench.err
