#!/bin/bash
# usage: tools/variant_bench.sh lib1.so lib2.so ...   -- short bench of library variants (MMGPU_LIB), one line each
for lib in "$@"; do
  MMGPU_LIB=$PWD/$lib python bench.py --no-cpu-baseline --steps 100 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['roofline']
print('%-40s step %.4f ms  k_map %.4f  k_fir %.4f  e2e %.4f ms' % ('$lib', d['ms_per_step'], r['path']['k_map_ms'], r['kernel_ms'], d['e2e']['ms_per_step']))
"
done
