#!/bin/bash
# usage: tools/ncu_kernel.sh <out-name> <kernel-regex> <skip> -- <filter_profile args...>
# one `ncu --set full` capture (with source) of one kernel launch, after a plain run of the same command
name=$1; regex=$2; skip=$3; shift 4
python tools/filter_profile.py "$@" > gpurun_out/plain_$name.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/$name python tools/filter_profile.py "$@" > gpurun_out/ncu_$name.log 2>&1
tail -1 gpurun_out/plain_$name.log
