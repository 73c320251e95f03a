# stage B of the update-kernel restructure: unit test of the B-MN-major / bulk round trip form, learner parity, timeline, A/B timing
timeout 300 python -m pytest tests/test_gpu_tc_gemm.py -m gpu -q --timeout=120 -k "b_mn" 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_learner.py -m gpu -q --timeout=300 -x 2>&1 | tail -15
timeout 300 bash profiles/debug/gpu_tl2.sh 2>&1 | grep "timeline" | head -4
timeout 300 python profiles/debug/variant_bench.py 2>&1 | tail -1
timeout 300 python profiles/debug/variant_bench.py 2>&1 | tail -1
