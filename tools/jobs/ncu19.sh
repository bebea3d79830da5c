#!/bin/bash
# GPU job: ncu --set full of every layer class of the shipped conv kernel (final round-2 build), DRAM traffic of one step, launch list
mkdir -p gpurun_out/n19
SAMPLES=4 ONLY=sorted timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -o gpurun_out/n19/prof_final python tools/profile_conv.py > gpurun_out/n19/ncu.log 2>&1
grep level gpurun_out/n19/ncu.log | head -12
timeout 600 ncu --profile-from-start off -k regex:conv_tc --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/n19/conv_traffic.csv python tools/step_once.py > gpurun_out/n19/traffic.log 2>&1
python tools/conv_traffic.py gpurun_out/n19/conv_traffic.csv gpurun_out/n19/traffic.json; cat gpurun_out/n19/traffic.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/n19/launches_step.csv python tools/step_once.py > gpurun_out/n19/step.log 2>&1
python tools/launch_summary.py gpurun_out/n19/launches_step.csv 2>/dev/null | head -30
ls -la gpurun_out/n19
