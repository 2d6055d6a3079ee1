#!/bin/bash
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/step_launches.csv \
    python tools/profile_step.py > gpurun_out/profile_step.log 2>&1; echo "launch list exit $?"
