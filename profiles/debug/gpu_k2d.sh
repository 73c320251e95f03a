cd uav-wrf-les-ppo-lstm_b200
run() { PLUME_NVCC_EXTRA="$1" python build.py --force > /dev/null 2>&1; echo "== $1"; (cd .. && python profiles/profile_k2.py 1048576 && python profiles/profile_k2.py 262144 && python profiles/profile_k2.py 131072); }
run "-DPLUME_STEP_NO_PREFETCH"
run "-DPLUME_STEP_THREADS=256 -DPLUME_STEP_MIN_BLOCKS=4"
run "-DPLUME_STEP_THREADS=64 -DPLUME_STEP_MIN_BLOCKS=16"
run "-DPLUME_STEP_THREADS=128 -DPLUME_STEP_MIN_BLOCKS=6"
run ""
