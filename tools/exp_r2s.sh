#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_visual_bert_gpu.py -x -q -m gpu -s -k base 2>&1 | tail -5
timeout 600 python -m pytest tests/test_graph_gpu.py -x -q -m gpu -s 2>&1 | tail -40
timeout 600 python -m pytest tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -3
