# in-run A/B of the product library against variant builds (boxes differ by ~1 %: only numbers of one run compare)
for i in 1 2; do
  timeout 60 python profiles/debug/variant_bench.py 2>&1 | tail -1
  for v in "$@"; do timeout 60 python profiles/debug/variant_bench.py profiles/debug/libplume_b200_$v.so 2>&1 | tail -1; done
done
