#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_solve.py -q -m gpu -x -s -k "large_ or unit_ids or beyond_the_supported or edge_cases" > gpurun_out/pytest_x.log 2>&1; echo "pytest exit $?"
tail -25 gpurun_out/pytest_x.log
