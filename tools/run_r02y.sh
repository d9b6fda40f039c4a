#!/bin/bash
# full GPU suite (new: ShortestPathAttr with a user metric through the generic pairwise driver)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02y_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02y_pytest_gpu.log; tail -12 gpurun_out/r02y_pytest_gpu.log | cut -c1-400
