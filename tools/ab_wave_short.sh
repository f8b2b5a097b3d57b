#!/bin/bash
# A/B of tuning builds of the wavefront kernel, short output: LIBS="tools/lib_a.so tools/lib_b.so" tools/ab_wave_short.sh [spp]
for rep in 1 2; do for lib in ${LIBS:-tools/lib_*.so}; do echo "== $lib (rep $rep)"; RT_B200_LIB=$PWD/$lib python tools/wave_bench.py ${1:-16} wave 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print('  ', d['name'], d['kw'], d['msamples_s'])"; done; done
