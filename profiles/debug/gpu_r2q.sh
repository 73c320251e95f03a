timeout 900 python -m pytest tests/test_gpu_learner.py tests/test_gpu_tc_gemm.py -m gpu -q --timeout=600 2>&1 | tail -15
python profiles/debug/variant_bench.py 2>&1 | tail -1
python profiles/debug/variant_bench.py 2>&1 | tail -1
