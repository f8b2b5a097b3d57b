#!/bin/bash
# A/B of tuning builds (tools/lib_*.so) in ONE gpurun call: c1 / book3 / c3 linear and BVH, twice
for rep in 1 2; do for lib in tools/lib_*.so; do echo "== $lib (rep $rep)"; RT_B200_LIB=$PWD/$lib python tools/ppl_probe.py 2>&1 | grep -E "ppl 2|mode 2" | tr '\n' ';'; echo; RT_B200_LIB=$PWD/$lib python tools/bvh_bench.py 16 2>&1 | grep 'lane": 1' | grep 'false' | cut -c1-140; RT_B200_LIB=$PWD/$lib python tools/quick_bench.py 32 2>&1 | grep '"c3"' | sed -n 2p | cut -c1-210; done; done
