# overlapped rollout collection: tests, then in-run A/B against the single launch ("0")
timeout 300 python -m pytest tests/test_gpu_rollout.py -x -q 2>&1 | tail -3
vb() { timeout 60 python profiles/debug/variant_bench.py "$@" 2>&1 | tail -1; }
for i in 1 2; do
vb - 0
vb
done
