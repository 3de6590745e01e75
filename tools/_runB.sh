timeout 300 python -m pytest tests -m gpu -x -q -k "vae or gemm or mlp or layered" 2>&1 | tail -2
timeout 120 python tools/vae_fused_check.py 2>&1 | tail -4
timeout 100 python tools/gemm_ab.py mannequin2_b200/libmq_b200.so
