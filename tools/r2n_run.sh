#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity_small.py tests/test_gpu_full_size.py tests/test_gpu_squelch.py -x -q 2>&1 | tail -12
