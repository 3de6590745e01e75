timeout 300 python -m pytest tests -m gpu -x -q -k "scan or gae or ppo or reinforce or trajectory" 2>&1 | tail -3
for v in 0 1 0 1; do echo "== MQ_GAE_VARIANT=$v"; MQ_GAE_VARIANT=$v timeout 100 python tools/time_hbm_kernels.py 2>&1 | grep -i gae; done
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pack_jobs -c 4 --csv --log-file gpurun_out/v9_pack_fp16x2.csv python tools/gemm_one.py 4096 4096 4096 fp16x2 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pack_jobs -c 4 --csv --log-file gpurun_out/v9_pack_bf16.csv python tools/gemm_one.py 4096 4096 4096 bf16 > /dev/null 2>&1
