# final check of the committed state: whole GPU suite and smoke()
timeout 150 python -m pytest tests -m gpu -q -x --timeout=100 2>&1 | tail -3
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
