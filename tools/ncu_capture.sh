#!/bin/bash
# usage (on the GPU box, under gpurun): tools/ncu_capture.sh <label> <kernel-regex> <bench.py arguments ...>
# One `ncu --set full` capture of the first matching launch after warm-up; leaves only small text files in gpurun_out/
# (<label>.raw.csv: every metric of the launch, <label>.lines.txt: instructions / stall samples per CUDA source line).
label=$1; kre=$2; shift 2
mkdir -p gpurun_out
rep=/tmp/$label.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$kre" -s 3 -c 1 -f -o /tmp/$label python bench.py "$@" --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/$label.ncu.log 2>&1
ncu -i $rep --page raw --csv > gpurun_out/$label.raw.csv 2>/dev/null
ncu -i $rep --page source --print-source cuda,sass --csv > /tmp/$label.src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/$label.src.csv 70 > gpurun_out/$label.lines.txt 2>&1
rm -f $rep /tmp/$label.src.csv
