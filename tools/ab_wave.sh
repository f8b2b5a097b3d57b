#!/bin/bash
# A/B of tuning builds (tools/lib_*.so) of the wavefront BVH kernel in ONE gpurun call, twice
for rep in 1 2; do for lib in tools/lib_*.so; do echo "== $lib (rep $rep) RT_TIE_CELL=${RT_TIE_CELL}"; RT_B200_LIB=$PWD/$lib python tools/wave_bench.py ${1:-16} wave 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print('  ', d['name'], d['kw'], d['msamples_s'], 'self', d['self_resolved_frac'], 'nodes/cast', d['node_visits_per_cast'], 'exact/cast', d['exact_per_cast'])"; done; done
