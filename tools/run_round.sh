python bench.py --network c3d --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3d.json 2> gpurun_out/bench_c3d.err; echo "c3d rc=$?"
tail -3 gpurun_out/bench_c3d.err
