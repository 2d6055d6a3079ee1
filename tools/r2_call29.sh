#!/bin/bash
# the preservation (DOP) micro-batch semantics on the GPU + the prior-target test next to it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_flux_engine.py -m gpu -q -x -k "preservation or prior_prediction" > gpurun_out/r2_call29_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r2_call29_tests.log
