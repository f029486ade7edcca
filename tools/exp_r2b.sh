#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_base_gpu.py -q -m gpu -s > gpurun_out/pytest_base_r2b.txt 2>&1; echo "base rc $? : $(tail -1 gpurun_out/pytest_base_r2b.txt)"
grep -E "base config|hidden-state|vision:|c1:|odd:|pre-train L|eval split|rotated|adaptive|grad rel err|FAILED|Error|^E " gpurun_out/pytest_base_r2b.txt | head -80
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_base_gpu.py > gpurun_out/pytest_r2b.txt 2>&1; echo "rest rc $? : $(tail -1 gpurun_out/pytest_r2b.txt)"
grep -E "FAILED|Error|^E " gpurun_out/pytest_r2b.txt | head -20
