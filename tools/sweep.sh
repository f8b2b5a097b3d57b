#!/bin/bash
# occupancy sweep over tuning builds (tools/lib_*.so): prints Msamples/s per (lib, paths/lane, cull source)
for lib in tools/lib_*.so; do
  echo "== $lib"
  RT_B200_LIB=$PWD/$lib python tools/quick_bench.py 32 2>&1 | grep '"c3"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   R=%d smem=%d  %.1f Msamples/s  %.0f Gtests/s' % (d['kw']['paths_per_lane'], d['kw']['cull_smem'], d['msamples_s'], d['gtests_s']))"
done
