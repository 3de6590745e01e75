timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/vI_tests.log 2>&1; echo tests rc=$? >> gpurun_out/vI_tests.log
timeout 600 python bench.py > gpurun_out/vI_bench.json 2> gpurun_out/vI_bench.err; echo bench rc=$? >> gpurun_out/vI_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/vI_bench_ref.json 2> gpurun_out/vI_bench_ref.err; echo rc=$? >> gpurun_out/vI_bench_ref.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/vI_smoke.log 2>&1
tail -3 gpurun_out/vI_tests.log; tail -2 gpurun_out/vI_smoke.log
