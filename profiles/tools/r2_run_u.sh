#!/bin/bash
O=gpurun_out/u1; mkdir -p $O
timeout 600 python profiles/tools/profile_step.py 64 120 800 fp16 $O/kernel_time.md > $O/profile.log 2>&1
timeout 600 python profiles/tools/timeline_full.py 64 120 800 fp16 20 > $O/timeline.txt 2>&1
head -5 $O/timeline.txt
