#!/bin/bash
# round 2, call 37: last sanity of the tree as built (smoke, the prepare / sharded / format tests, a short C3 line)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_prepare.py tests/test_gpu_parity.py -x -q --timeout 300 --timeout-method thread > gpurun_out/r2al_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2al_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-ingest --no-cpu-baseline > gpurun_out/r2al_c3.log 2>&1; echo "rc=$?"
python tools/bline.py gpurun_out/r2al_c3.log
