timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | tail -1
