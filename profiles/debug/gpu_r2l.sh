timeout 600 python -m pytest tests/test_gpu_rollout.py -m gpu -q --timeout=300 -k "other_hidden" 2>&1 | tail -3
timeout 600 python profiles/sweep.py --hidden 128 256 --envs 4096 2>&1 | cut -c1-330
