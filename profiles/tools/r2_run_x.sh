#!/bin/bash
O=gpurun_out/${1:-x1}; mkdir -p $O
timeout 600 python profiles/tools/timeline_e2e.py 64 120 800 fp16 > $O/timeline_e2e.txt 2>&1
grep -v Warn $O/timeline_e2e.txt | head -120
