timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/vH_multi_gpu_pytest.log 2>&1; echo rc=$? >> gpurun_out/vH_multi_gpu_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/vH_bench_2gpu.json 2> gpurun_out/vH_bench_2gpu.err; echo rc=$? >> gpurun_out/vH_bench_2gpu.err
tail -3 gpurun_out/vH_multi_gpu_pytest.log
