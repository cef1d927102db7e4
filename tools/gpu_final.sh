#!/bin/bash
# state refresh of the FINAL build: GPU suite, smoke, traffic + tensor-pipe capture (fingerprinted), bench with all extras,
# reference arm, layer table, launch list, C3D launch list, ncu --set full of the top kernels
T=${1:-r02_final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
timeout 300 python bench.py --quick --no-graph --steps 1 --warmup 1 --no-extras --no-cpu-baseline > /dev/null 2>&1 && \
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:igemm --csv --log-file gpurun_out/${T}_conv_traffic.csv python bench.py --quick --no-graph --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/${T}_traffic_ncu.log 2>&1
python tools/conv_traffic.py gpurun_out/${T}_conv_traffic.csv profiles/r02_conv_dram_traffic.json | cut -c1-300
cp profiles/r02_conv_dram_traffic.json gpurun_out/${T}_conv_dram_traffic.json
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-200 gpurun_out/${T}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; cut -c1-300 gpurun_out/${T}_bench_reference.json
timeout 600 python bench.py --no-cpu-baseline --no-extras --layer-table > /dev/null 2> gpurun_out/${T}_layer_table.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --quick --no-graph --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/${T}_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/${T}_launches.csv > gpurun_out/${T}_launch_summary.txt; head -16 gpurun_out/${T}_launch_summary.txt
timeout 300 python bench.py --network c3d --quick --no-graph --steps 2 --warmup 1 --no-extras --no-cpu-baseline > /dev/null 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_c3d_launches.csv python bench.py --network c3d --quick --no-graph --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/${T}_c3d_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/${T}_c3d_launches.csv > gpurun_out/${T}_c3d_launch_summary.txt; head -8 gpurun_out/${T}_c3d_launch_summary.txt
cap() { # name spec pass skip count
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm -s $4 -c $5 -f -o gpurun_out/${T}_$1 python tools/ncu_one.py $2 $3 > gpurun_out/${T}_$1.log 2>&1; echo "$1 rc=$?"; }
cap fprop_64_144 22,16,56,56,64,144,1,3,3,1,1,1,0,1,1 fprop 2 1
cap dgrad_fused_64_144 22,16,56,56,64,144,1,3,3,1,1,1,0,1,1 dgrad_fused 2 1
cap fprop_144_64 22,16,56,56,144,64,3,1,1,1,1,1,1,0,0 fprop 2 1
cap dgrad_144_64 22,16,56,56,144,64,3,1,1,1,1,1,1,0,0 dgrad 2 1
cap fprop_128_288 22,8,28,28,128,288,1,3,3,1,1,1,0,1,1 fprop 2 1
