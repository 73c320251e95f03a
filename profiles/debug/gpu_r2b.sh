set -x
python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu_r2b.log 2>&1; tail -40 gpurun_out/pytest_gpu_r2b.log
