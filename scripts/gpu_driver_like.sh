#!/bin/bash
# What the driver runs at round end, in one process each: the whole GPU suite, smoke(), then the launch list of one step.
mkdir -p gpurun_out
o=gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu ) > $o/r02_gpu_suite_one_process.txt 2>&1; tail -4 $o/r02_gpu_suite_one_process.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/r02_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-decode > $o/r02_launches_bench.log 2>&1
ls -la $o/r02_launches.csv
