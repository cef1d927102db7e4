#!/bin/bash
# final-state check: full GPU suite, smoke, default bench, FusedAdam A/B
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r9_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r9_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r9_smoke.log
python bench.py > gpurun_out/r9_bench.json 2> gpurun_out/r9_bench.err
python bench.py --no-cpu-baseline --optimizer zsv > gpurun_out/r9_bench_zsvadam.json 2> gpurun_out/r9_bench_zsvadam.err
python bench.py --no-cpu-baseline > gpurun_out/r9_bench_b.json 2> gpurun_out/r9_bench_b.err
tail -3 gpurun_out/r9_pytest.log; tail -2 gpurun_out/r9_smoke.log
