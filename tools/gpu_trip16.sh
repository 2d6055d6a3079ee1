#!/bin/bash
timeout 400 python -X faulthandler -m pytest tests/test_gpu_attention.py -x -q --timeout 200 --timeout-method=thread -k "not timing" 2>&1 | tail -3
B200_ATTN_FWD=1 timeout 120 python tools/time_attn.py 2>&1 | head -1
B200_ATTN_FWD=2 timeout 120 python tools/time_attn.py 2>&1 | head -1
