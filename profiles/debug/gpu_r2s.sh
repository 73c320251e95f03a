# update-kernel restructure: learner parity, run-to-run probe, timeline, A/B timing (tight timeouts: a hung kernel must not eat the GPU budget)
timeout 150 python -m pytest tests/test_gpu_learner.py -m gpu -q --timeout=100 2>&1 | tail -2
timeout 100 python profiles/debug/flaky_probe.py 2>&1 | tail -5
timeout 60 bash profiles/debug/gpu_tl2.sh 2>&1 | grep "timeline\|issuer" | head -3
timeout 60 python profiles/debug/variant_bench.py 2>&1 | tail -1
timeout 60 python profiles/debug/variant_bench.py 2>&1 | tail -1
for v in "$@"; do timeout 60 python profiles/debug/variant_bench.py profiles/debug/libplume_b200_$v.so 2>&1 | tail -1; done
