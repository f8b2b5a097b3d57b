# final 1-GPU evidence pass: GPU tests, bench.py (both arms), ncu captures of the two render kernels (each after the same
# command exited 0 without ncu), launch list of a short bench run
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2z_tests.log; cat gpurun_out/r2z_tests.log
python bench.py > gpurun_out/r2z_bench_1gpu.json 2> gpurun_out/r2z_bench_1gpu.err; tail -2 gpurun_out/r2z_bench_1gpu.err; head -c 600 gpurun_out/r2z_bench_1gpu.json; echo
python bench.py --impl reference > gpurun_out/r2z_bench_ref_1gpu.json 2> gpurun_out/r2z_bench_ref_1gpu.err; head -c 400 gpurun_out/r2z_bench_ref_1gpu.json; echo
NCU="ncu --set full --clock-control none --import-source on -s 1 -c 1 -f"
python tools/profile_render.py 8 > gpurun_out/r2z_plain_scan.log 2>&1 && $NCU -k regex:render_kernel -o gpurun_out/r2z_scan_c3 python tools/profile_render.py 8 > gpurun_out/r2z_ncu_scan.log 2>&1; tail -2 gpurun_out/r2z_ncu_scan.log
python tools/profile_render.py 32 0 1200 800 0 0 3 11 > gpurun_out/r2z_plain_wave_c3.log 2>&1 && $NCU -k regex:render_wave -o gpurun_out/r2z_wave_c3 python tools/profile_render.py 32 0 1200 800 0 0 3 11 > gpurun_out/r2z_ncu_wave_c3.log 2>&1; tail -2 gpurun_out/r2z_ncu_wave_c3.log
python tools/profile_render.py 16 0 1920 1080 0 0 3 158 > gpurun_out/r2z_plain_wave_c4.log 2>&1 && $NCU -k regex:render_wave -o gpurun_out/r2z_wave_c4 python tools/profile_render.py 16 0 1920 1080 0 0 3 158 > gpurun_out/r2z_ncu_wave_c4.log 2>&1; tail -2 gpurun_out/r2z_ncu_wave_c4.log
python bench.py --steps 2 --warmup 1 --spp 16 --no-other-configs --no-cpu-baseline > gpurun_out/r2z_bench_small.json 2> gpurun_out/r2z_bench_small.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 1 --spp 16 --no-other-configs --no-cpu-baseline > gpurun_out/r2z_ncu_launch.log 2>&1; tail -2 gpurun_out/r2z_ncu_launch.log
