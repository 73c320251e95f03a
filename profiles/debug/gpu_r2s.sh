# update-kernel restructure: learner parity, timeline, A/B timing
timeout 900 python -m pytest tests/test_gpu_learner.py -m gpu -q --timeout=300 -x 2>&1 | tail -15
timeout 300 bash profiles/debug/gpu_tl2.sh 2>&1 | grep "timeline" | head -3
timeout 300 python profiles/debug/variant_bench.py 2>&1 | tail -1
timeout 300 python profiles/debug/variant_bench.py 2>&1 | tail -1
