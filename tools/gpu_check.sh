#!/bin/bash
# quick check: the GPU test suite + smoke
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/check_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/check_pytest.log; tail -8 gpurun_out/check_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
