#!/bin/bash
# quick GPU sanity run used during development (under gpurun)
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 240 python -m pytest tests/test_gpu_als.py -q -m gpu -x --timeout 60 2>&1 | tail -30
timeout 300 python bench.py --steps 40 --warmup 5 2>&1 | tail -5
