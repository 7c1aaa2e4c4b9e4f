#!/bin/bash
# round 2, run 63: last full GPU suite + smoke on the final tree
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
