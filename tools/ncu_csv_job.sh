#!/bin/bash
# ncu --set full of kernels matching a regex, exported on the box as the reduced CSV (the .ncu-rep with sources can exceed
# gpurun's 64 MiB return limit).  usage: bash tools/ncu_csv_job.sh <tag> <count> <regex> <cmd...>  -> gpurun_out/ncu_<tag>.csv
set -u
tag=$1; cnt=$2; rx=$3; shift 3
mkdir -p gpurun_out
timeout 600 "$@" > gpurun_out/ncu_${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -3 gpurun_out/ncu_${tag}_plain.log; exit 1; }
timeout 1500 ncu --set full --clock-control none -k regex:"$rx" -c "$cnt" -f -o /tmp/ncu_$tag "$@" > /tmp/ncu_$tag.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/ncu_$tag.ncu-rep --page raw --csv > /tmp/ncu_${tag}_raw.csv 2>/dev/null
python tools/ncu_family_columns.py /tmp/ncu_${tag}_raw.csv gpurun_out/ncu_$tag.csv
wc -l gpurun_out/ncu_$tag.csv
