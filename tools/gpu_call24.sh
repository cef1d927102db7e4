#!/bin/bash
mkdir -p gpurun_out
cap() { # name spec pass skip count
timeout 300 python tools/ncu_one.py $2 $3 > /dev/null 2>&1 || { echo "$1 plain run failed"; return; }
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm -s $4 -c $5 -f -o gpurun_out/c24_$1 python tools/ncu_one.py $2 $3 > gpurun_out/c24_$1.log 2>&1; echo "$1 rc=$?"; }
cap fprop_144_64 22,16,56,56,144,64,3,1,1,1,1,1,1,0,0 fprop 2 1
