#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu8.log
tail -15 gpurun_out/pytest_gpu8.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
