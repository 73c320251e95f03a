# is the spread of the overlapped rollout the race between the first head and the second lockstep launch?
vb() { timeout 60 python profiles/debug/variant_bench.py "$@" 2>&1 | tail -1; }
for i in 1 2; do
vb profiles/debug/libplume_b200_zp.so 32,224 0
vb profiles/debug/libplume_b200_zp.so 32,224 4
vb profiles/debug/libplume_b200_zp.so 32,224 10
done
