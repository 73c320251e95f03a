for v in "$@"; do python profiles/debug/variant_bench.py profiles/debug/libplume_b200_$v.so 2>&1 | tail -1; done
