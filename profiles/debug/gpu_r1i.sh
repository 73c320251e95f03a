python -m pytest tests/test_gpu_lstm_train.py tests/test_gpu_rollout.py tests/test_gpu_env.py -x -q 2>&1 | tail -3
python profiles/profile_lstm_train.py 64 && python profiles/profile_lstm_train.py 512
python bench.py --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_r1i.log 2> gpurun_out/bench_r1i.err; tail -c 400 gpurun_out/bench_r1i.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r1i.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['rollout_env_steps_per_sec'], d['kernels']['rollout']); print(d['lstm_train']); print({k:(v['ms'],round(v['frac'],3)) for k,v in d['plume_kernels'].items()})"
