#!/bin/bash
for rep in 1 2; do for lib in ${LIBS:-tools/lib_*.so}; do echo "== $lib (rep $rep)"; RT_B200_LIB=$PWD/$lib python tools/scan_bench.py ${1:-32} 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print('  ', d['name'], d['kw'], d['msamples_s'], 'frac', d['frac'], 'exact/cast', d['exact_per_cast'])"; done; done
