#!/bin/bash
# Round check on the GPU box: parity tests, headline bench (both arms), launch list, full ncu captures.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 200 --warmup 5 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1500 gpurun_out/bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 4 --warmup 3 --no-also --no-cpu-baseline --e2e-steps 1 --graph-steps 0 > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_hd_warp -s 4 -c 2 -o gpurun_out/prof_warp9 -f \
   python scripts/run_cfg.py formation_hd_env 9 131072 8 > gpurun_out/ncu_warp9.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 3 -c 1 -o gpurun_out/prof_tile243 -f \
   python scripts/run_cfg.py formation_hd_env 243 1024 5 > gpurun_out/ncu_tile243.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
ls -la gpurun_out
