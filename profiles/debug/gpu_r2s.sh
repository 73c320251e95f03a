# update-kernel restructure: learner parity, timeline, A/B timing
timeout 900 python -m pytest tests/test_gpu_learner.py -m gpu -q --timeout=300 -x 2>&1 | tail -15
timeout 300 bash profiles/debug/gpu_tl2.sh 2>&1 | grep "timeline\|issuer" | head -3
timeout 300 python profiles/debug/variant_bench.py 2>&1 | tail -1
timeout 300 python profiles/debug/variant_bench.py 2>&1 | tail -1
for v in "$@"; do timeout 300 python profiles/debug/variant_bench.py profiles/debug/libplume_b200_$v.so 2>&1 | tail -1; done
