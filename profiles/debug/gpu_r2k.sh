timeout 600 python -m pytest tests/test_gpu_rollout.py -m gpu -q --timeout=300 -k "other_hidden" > gpurun_out/pytest_gpu_r2k.log 2>&1; tail -25 gpurun_out/pytest_gpu_r2k.log
timeout 600 python profiles/sweep.py --hidden 64 128 256 --envs 4096 > gpurun_out/sweep_r2k.jsonl 2> gpurun_out/sweep_r2k.err; cat gpurun_out/sweep_r2k.jsonl; tail -3 gpurun_out/sweep_r2k.err
