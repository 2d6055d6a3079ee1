#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 2 -c 6 -o gpurun_out/r1c_attn -f \
    python tools/profile_kernels.py attn > gpurun_out/prof_attn.log 2>&1; echo "attn full exit $?"
