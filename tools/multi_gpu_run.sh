# 8-GPU pass (one node): multi-device tests, tiles vs samples deal with per-rank kernel times, bench.py --gpus 8, rt_main --gpus 8
set -x
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi or two_devices or sharded or sample_split" 2>&1 | tail -4 > gpurun_out/r2z_t8_multi.log; cat gpurun_out/r2z_t8_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/deal_bench.py c3 > gpurun_out/r2z_deal_8gpu.log 2> gpurun_out/r2z_deal_8gpu.err; cat gpurun_out/r2z_deal_8gpu.log; tail -3 gpurun_out/r2z_deal_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2z_bench_8gpu.json 2> gpurun_out/r2z_bench_8gpu.err; head -c 400 gpurun_out/r2z_bench_8gpu.json; tail -3 gpurun_out/r2z_bench_8gpu.err
( time ./petershirleyraytracer_b200/rt_main 1200 100 50 1 --gpus 8 > /tmp/m8.ppm ) 2> gpurun_out/r2z_rtmain_8gpu.log; ( time ./petershirleyraytracer_b200/rt_main 1200 100 50 1 > /tmp/m1.ppm ) 2>> gpurun_out/r2z_rtmain_8gpu.log; cmp /tmp/m1.ppm /tmp/m8.ppm && echo "rt_main --gpus 8 == 1 GPU" >> gpurun_out/r2z_rtmain_8gpu.log; cat gpurun_out/r2z_rtmain_8gpu.log
