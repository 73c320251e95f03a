timeout 600 python -m pytest tests/test_gpu_env.py -m gpu -q --timeout=300 2>&1 | tail -3
python - <<'PY'
import sys, json, torch
sys.path.insert(0,'.')
import bench
import uav_wrf_les_ppo_lstm_b200 as pb
r = bench.plume_kernel_rooflines(pb, torch, bench.measured_peaks(), torch.device("cuda:0"))
for k,v in r.items(): print(k, round(v['ms'],5), round(v['frac'],4))
PY
