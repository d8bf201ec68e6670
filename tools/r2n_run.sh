#!/bin/bash
timeout 300 python tools/rem_probe.py 2>&1 | tail -25
