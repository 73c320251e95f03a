timeout 900 python -m pytest tests/test_gpu_evaluate.py tests/test_gpu_env.py -m gpu -q --timeout=600 2>&1 | tail -25
python - <<'PY'
import sys, json, torch, time
sys.path.insert(0,'.')
import bench
import uav_wrf_les_ppo_lstm_b200 as pb
r = bench.plume_kernel_rooflines(pb, torch, bench.measured_peaks(), torch.device("cuda:0"))
for k,v in r.items(): print(k, round(v['ms'],5), round(v['frac'],4))
torch.manual_seed(0)
model = pb.PPOActorCritic(device="cuda")
for stop, kw in (("lstm", {"head": pb.PeakAndStopPredictor(device="cuda")}), ("threshold", {"head": pb.ConcentrationThresholdPredictor(device="cuda"), "scaler": (0.0, 100.0)}), ("fixed", {})):
    for rep in range(2):
        t0=time.perf_counter(); res = pb.evaluate_policy(model, stop=stop, num_envs=4096, seed=1, **kw); torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print(stop, "episodes/s", 4096/dt, "env-steps/s", float(res.steps.double().sum())/dt, "mean steps", float(res.steps.double().mean()), res.summary())
PY
