#!/bin/bash
# Round evidence on the GPU box: ncu --set full captures of every named kernel/config (exported to CSV on the box),
# the ncu launch list of the bench command, the bench lines of both arms.  Usage: scripts/round_profiles.sh TAG
TAG=${1:-r02}
mkdir -p gpurun_out
bash scripts/gpu_prof.sh $TAG \
  "formation_hd_env 9 131072 1 k_hd_warp all" "formation_hd_env 9 131072 1 k_hd_warp none" \
  "formation_hd_env 9 1048576 1 k_hd_warp none" "formation_hd_env 27 65536 1 k_hd_warp none" \
  "formation_hd_env 3 1048576 1 k_hd_warp none" "basic_formation_env 3 1048576 1 k_hd_warp none" \
  "formation_hd_env 4 262144 1 k_hd_warp none" \
  "formation_hd_env 243 1024 1 k_step none" "formation_hd_env 243 1024 0 k_step none" \
  "formation_hd_env 243 8192 0 k_step none" "formation_hd_env 81 8192 1 k_step none" \
  "formation_hd_partial_env 4 262144 1 k_lm_warp none" "formation_hd_partial_env 5 262144 1 k_lm_warp none" \
  "formation_hd_obs_env 4 262144 1 k_lm_warp none" > gpurun_out/${TAG}_prof.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_policy -s 2 -c 1 -o gpurun_out/${TAG}_policy9 -f \
  python scripts/run_policy.py 9 131072 > /dev/null 2>&1
ncu -i gpurun_out/${TAG}_policy9.ncu-rep --page raw --csv > gpurun_out/${TAG}_policy9.raw.csv
ncu -i gpurun_out/${TAG}_policy9.ncu-rep --page source --csv | gzip > gpurun_out/${TAG}_policy9.source.csv.gz
rm -f gpurun_out/${TAG}_policy9.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 4 --warmup 3 --no-also --no-cpu-baseline --no-strong --e2e-steps 1 --graph-steps 0 > gpurun_out/${TAG}_bench_ncu.log 2>&1
( time timeout 300 python bench.py --impl reference --steps 100 --warmup 5 ) > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
tail -3 gpurun_out/${TAG}_bench_n1.err
ls -la gpurun_out | tail -40
