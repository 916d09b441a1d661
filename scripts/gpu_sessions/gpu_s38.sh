#!/bin/bash
# the driver's own round-end commands on the final build
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
