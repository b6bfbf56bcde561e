#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_multigrid.py -m gpu -x -q -s ) > gpurun_out/pytest_mg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mg.log; tail -40 gpurun_out/pytest_mg.log
