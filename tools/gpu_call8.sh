#!/bin/bash
mkdir -p gpurun_out
VDQN_NVCC_FLAGS=-DVDQN_ROLE_PROFILE python video_dqn_b200/build.py --force > /dev/null 2>&1
timeout 300 python tools/role_profile_pool.py > gpurun_out/c8_role_pool.txt 2>&1
cat gpurun_out/c8_role_pool.txt
