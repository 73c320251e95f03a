python profiles/debug/variant_bench.py profiles/debug/libplume_b200_rtl.so 2>&1 | grep -E "timeline|iteration" | head -40
