#!/bin/bash
mkdir -p gpurun_out
cap() { # name spec pass count
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm -s $4 -c $5 -f -o gpurun_out/c19_$1 python tools/ncu_one.py $2 $3 > gpurun_out/c19_$1.log 2>&1; echo "$1 rc=$?"; }
cap dgrad_45_64 22,16,56,56,45,64,3,1,1,1,1,1,1,0,0 dgrad 1 1
cap fprop_stem 22,16,112,112,3,45,1,7,7,1,2,2,0,3,3 fprop 1 1
cap dgrad_288_128 22,8,28,28,288,128,3,1,1,1,1,1,1,0,0 dgrad 1 1
cap dgrad_64_230 22,16,56,56,64,230,1,3,3,1,2,2,0,1,1 dgrad 4 4
