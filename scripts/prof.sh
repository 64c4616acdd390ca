#!/bin/bash
# usage (under gpurun): scripts/prof.sh <kernel regex> <name> <skip> <count> <cmd...>   -> gpurun_out/<name>.ncu-rep + log
k=$1; name=$2; skip=$3; cnt=$4; shift 4
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c $cnt -o gpurun_out/$name -f "$@" > gpurun_out/$name.log 2>&1
tail -4 gpurun_out/$name.log
