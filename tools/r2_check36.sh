#!/bin/bash
# round 2, call 36: the whole GPU suite and smoke() on the final tree
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2aj_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2aj_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
