#!/bin/bash
mkdir -p gpurun_out/t
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t/tests.log 2>&1
tail -15 gpurun_out/t/tests.log
