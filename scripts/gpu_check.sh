#!/bin/bash
# Round check on the GPU box: parity tests, headline bench (both arms), launch list.  Usage: scripts/gpu_check.sh [tag]
TAG=${1:-r02}
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -15 gpurun_out/${TAG}_pytest_gpu.log
( time timeout 300 python bench.py --impl reference --steps 100 --warmup 5 ) > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; cat gpurun_out/${TAG}_bench_reference.json
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -c 3000 gpurun_out/${TAG}_bench_n1.json; tail -5 gpurun_out/${TAG}_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 4 --warmup 3 --no-also --no-cpu-baseline --no-strong --e2e-steps 1 --graph-steps 0 > gpurun_out/${TAG}_bench_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
ls -la gpurun_out
