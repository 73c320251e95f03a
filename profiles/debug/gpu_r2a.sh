set -x
python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu_r2a.log 2>&1; tail -15 gpurun_out/pytest_gpu_r2a.log
python profiles/debug/variant_bench.py 2>&1 | tail -1
python profiles/debug/variant_bench.py profiles/debug/libplume_b200_noqbar.so 2>&1 | tail -1
python profiles/debug/variant_bench.py 2>&1 | tail -1
python bench.py --skip-cpu > gpurun_out/bench_r2a.log 2> gpurun_out/bench_r2a.err; tail -c 600 gpurun_out/bench_r2a.err; tail -c 1500 gpurun_out/bench_r2a.log
