#!/bin/bash
# tools/exp_run.sh -- on the GPU box: parity of the product build, then Tx timings; all under `timeout`
timeout 60 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
