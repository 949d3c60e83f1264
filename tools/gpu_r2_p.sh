#!/bin/bash
set -u
export PYTHONPATH=.
timeout 600 python tools/bench_device_dataset.py 1024 2>&1 | tail -2
