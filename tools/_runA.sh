timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/vA_tests.log 2>&1; echo tests rc=$? >> gpurun_out/vA_tests.log
timeout 600 python bench.py > gpurun_out/vA_bench.json 2> gpurun_out/vA_bench.err; echo bench rc=$? >> gpurun_out/vA_bench.err
tail -3 gpurun_out/vA_tests.log
