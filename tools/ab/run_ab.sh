#!/bin/bash
# A/B of two builds of libhello_moe.so on one GPU box: tools/ab/run_ab.sh <sites> <lib_a> <lib_b> ...
sites=$1; shift
for rep in 1 2; do
  for lib in "$@"; do
    HELLO_MOE_LIB=$PWD/$lib python bench.py --sites $sites --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('%-40s %9.1f k sites/s  step %.1f ms  rc stage %.1f ms  frac %.4f  clk %s' % ('$lib', d['value'] / 1e3, d['ms_per_step'], d['roofline']['stage_ms_per_step'], d['roofline']['frac'], d['clocks'].get('sm_mhz')))"
  done
done
