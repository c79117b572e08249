#!/bin/bash
# A/B of builds of libhello_moe.so on one GPU box for any workload: tools/ab/run_ab_wl.sh <workload> <sites> <lib_a> <lib_b> ...
wl=$1; sites=$2; shift; shift
for rep in 1 2; do
  for lib in "$@"; do
    HELLO_MOE_LIB=$PWD/$lib python bench.py --workload $wl --sites $sites --steps 3 --warmup 1 --no-e2e --no-cpu-baseline --no-other-workloads 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('%-28s %9.1f k sites/s  step %.1f ms  clk %s' % ('$lib', d['value'] / 1e3, d['ms_per_step'], d['clocks'].get('sm_mhz')))"
  done
done
