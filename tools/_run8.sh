timeout 100 python tools/gemm_one.py 4096 4096 4096 fp16x2 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -s 2 -c 1 -o gpurun_out/r2_gemm_4096cubed_fp16x2 -f python tools/gemm_one.py 4096 4096 4096 fp16x2 > gpurun_out/v8_ncu1.log 2>&1
timeout 100 python tools/gemm_one.py 4096 4096 4096 bf16 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -s 2 -c 1 -o gpurun_out/r2_gemm_4096cubed_bf16 -f python tools/gemm_one.py 4096 4096 4096 bf16 > gpurun_out/v8_ncu2.log 2>&1
timeout 100 python tools/vae_launches.py && timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -s 25 -c 1 -o gpurun_out/r2_gemm_vae_dec_heads -f python tools/vae_launches.py > gpurun_out/v8_ncu3.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v8_vae_launches.csv python tools/vae_launches.py > /dev/null 2>&1
timeout 300 python bench.py --workload vae --no-cpu-baseline > gpurun_out/v8_bench_vae.json 2> gpurun_out/v8_bench_vae.err
ls -la gpurun_out/*.ncu-rep | tail -4
