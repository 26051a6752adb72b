#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_device_records.py -m gpu -q -x ) > gpurun_out/r02_c8_pytest_records.log 2>&1
tail -30 gpurun_out/r02_c8_pytest_records.log
