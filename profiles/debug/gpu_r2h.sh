set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_equivalence.py > gpurun_out/r2_multi_gpu_equivalence_n2.log 2>&1; tail -6 gpurun_out/r2_multi_gpu_equivalence_n2.log
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_multi_r2h.log 2>&1; tail -3 gpurun_out/pytest_multi_r2h.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --skip-cpu --skip-aux > gpurun_out/bench_r2h_n2.log 2> gpurun_out/bench_r2h_n2.err; tail -c 400 gpurun_out/bench_r2h_n2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2h_n2.log').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['rollout_env_steps_per_sec'], d.get('rank_skew'))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 profiles/debug/n_gpu_breakdown.py 2>&1 | tail -6
