#!/bin/bash
# last call of round 2: the full GPU suite on the final tree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 110 python -m pytest tests -m gpu -q > gpurun_out/s2j_pytest.log 2>&1
tail -3 gpurun_out/s2j_pytest.log
