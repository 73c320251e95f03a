# stop heads after the warp-uniform issuer change: parity tests, the iteration, the hidden-size sweep
timeout 200 python -m pytest tests/test_gpu_policy.py tests/test_gpu_rollout.py tests/test_gpu_evaluate.py -m gpu -q --timeout=150 2>&1 | tail -3
timeout 60 python profiles/debug/variant_bench.py 2>&1 | tail -1
timeout 60 python profiles/debug/variant_bench.py 2>&1 | tail -1
timeout 200 python profiles/sweep.py --envs 4096 --hidden 32 64 128 256 2>/dev/null | tail -6 | cut -c1-330
