#!/bin/bash
mkdir -p gpurun_out
bash scripts/ncu_launch_list.sh
bash scripts/ncu_kernel_captures.sh
