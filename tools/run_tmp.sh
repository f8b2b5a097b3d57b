set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2z_tests.log; cat gpurun_out/r2z_tests.log
python bench.py > gpurun_out/r2z_bench_1gpu.json 2> gpurun_out/r2z_bench_1gpu.err; tail -2 gpurun_out/r2z_bench_1gpu.err; head -c 300 gpurun_out/r2z_bench_1gpu.json; echo
NCU="ncu --set full --clock-control none --import-source on -s 1 -c 1 -f"
python tools/profile_render.py 8 > gpurun_out/r2z_plain_scan.log 2>&1 && $NCU -k regex:render_kernel -o gpurun_out/r2z_scan_c3 python tools/profile_render.py 8 > gpurun_out/r2z_ncu_scan.log 2>&1; tail -2 gpurun_out/r2z_ncu_scan.log
