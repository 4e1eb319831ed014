#!/bin/bash
# helper used with gpurun: run the GPU tests and keep logs in gpurun_out/
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | tail -80 | tee gpurun_out/pytest_gpu.log
