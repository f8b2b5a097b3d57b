#!/bin/bash
# A/B of tuning builds (tools/lib_*.so) in ONE gpurun call: c1 / book3 / c3 linear and BVH, twice
for rep in 1 2; do for lib in tools/lib_*.so; do echo "== $lib (rep $rep)"; RT_B200_LIB=$PWD/$lib python tools/ppl_probe.py 2>&1 | grep -E "ppl 2|mode 2" | awk '{printf "%s m%s %s; ", $1, $3, $6}'; echo; RT_B200_LIB=$PWD/$lib python tools/bvh_bench.py 16 2>&1 | grep 'lane": 1' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('   bvh', d['name'], 'eo' if d['kw']['early_out'] else '  ', d['msamples_s'], 'boxes/cast', d['node_tests_per_cast'], 'exact/cast', d['exact_per_cast'])"; done; done
