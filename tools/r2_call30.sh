#!/bin/bash
# last sanity of the committed tree: the whole GPU tier
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_last_gputests.log 2>&1; echo "gpu tests exit $?"; tail -3 gpurun_out/r2_last_gputests.log
