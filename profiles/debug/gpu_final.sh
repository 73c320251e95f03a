# final check of the committed state: whole GPU suite, smoke(), the default bench line
timeout 600 python -m pytest tests -m gpu -q --timeout=200 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
tail -c 300 gpurun_out/r2f_bench_n1.json; echo
