"""K2 (plume step) alone at 2^20 envs, procedural field, for ncu: plain run first, then the same command
under `ncu --set full -k regex:step_kernel`.  argv[1] = envs, argv[2] = 'fast' for PLUME_FLAG_FAST_REWARD."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uav_wrf_les_ppo_lstm_b200 as pb  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    fast = len(sys.argv) > 2 and sys.argv[2] == "fast"
    if os.environ.get("PLUME_L2_FETCH"):     # experiment: cudaLimitMaxL2FetchGranularity (0x05)
        import ctypes
        rt = ctypes.CDLL("libcudart.so.12")
        torch.cuda.init(); torch.zeros(1, device="cuda")
        print("cudaDeviceSetLimit rc", rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["PLUME_L2_FETCH"]))))
    env = pb.VecMethaneEnv(n, field_mode="procedural", auto_reset=True, seed=2, fast_reward=fast)
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(int(os.environ.get("PLUME_K2_WALK", "30"))):      # walk away from the corner so the profiled step sees a typical state mix
        env.step(torch.randint(0, 5, (n,), dtype=torch.int32, device="cuda", generator=g))
    torch.cuda.synchronize()
    import ctypes as C
    lib = pb._lib.load()
    acts = torch.randint(0, 5, (n,), dtype=torch.int32, device="cuda", generator=g)
    flags = pb._lib.FLAG_AUTO_RESET | (pb._lib.FLAG_FAST_REWARD if fast else 0)
    stream = torch.cuda.current_stream().cuda_stream

    def launch(reps):      # straight through the C ABI: no per-call Python tensor handling in the timed region
        for _ in range(reps):
            lib.plume_env_step(C.byref(env.c_config), C.byref(env.c_state), acts.data_ptr(), None, flags,
                               env.obs.data_ptr(), env.reward.data_ptr(), env.done.data_ptr(), env.reached.data_ptr(),
                               env.info_t.data_ptr(), env.final_obs.data_ptr(), None, stream)
    launch(5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch(20)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"profile_k2 n={n} fast={fast} {ms:.4f} ms/step {n / ms * 1e3:.3e} env-steps/s {n * 146 / ms / 1e6:.0f} GB/s (146 B/env-step)")


if __name__ == "__main__":
    main()
