#!/bin/bash
# per-kernel times (ncu launch list) of the head kernels for the given builds: tools/ab/run_heads.sh <lib> ...
for lib in "$@"; do
  HELLO_MOE_LIB=$PWD/$lib ncu --metrics gpu__time_duration.sum --clock-control none -k regex:headconv_tc --csv --log-file /tmp/h.csv python bench.py --sites 16384 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-other-workloads > /dev/null 2>&1
  echo "== $lib"; python tools/launch_shares.py /tmp/h.csv 4
done
