#!/bin/bash
# round 2, call 26: the command line at BASELINE C3 size (small rehearsal first)
mkdir -p gpurun_out
df -h /dev/shm /tmp | tail -3
free -g | head -2
timeout 300 python tools/cli_fullsize.py 50000000 1000000 > gpurun_out/r2_cli_rehearsal.txt 2>&1; echo "rehearsal rc=$?"; tail -12 gpurun_out/r2_cli_rehearsal.txt
timeout 1500 python tools/cli_fullsize.py > gpurun_out/r2_cli_fullsize.txt 2>&1; echo "fullsize rc=$?"; tail -40 gpurun_out/r2_cli_fullsize.txt
