#!/bin/bash
# ncu --set full captures of the fused step kernel for a list of configurations (one GPU).  The .ncu-rep files are
# ~23 MB each (gpurun_out is capped at 64 MiB), so each report is exported on the box to a raw-page CSV (all metrics)
# and a source-page CSV (per-instruction stall samples; needs -lineinfo) and then deleted unless KEEP_REP=1.
# Usage: scripts/gpu_prof.sh TAG "SCEN N E OBS KERNELREGEX [cache]" ...   (cache = all|none: ncu --cache-control)
TAG=$1; shift
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg
  SCEN=$1; N=$2; E=$3; OBS=$4; KRE=$5; CC=${6:-all}
  NAME=${TAG}_${SCEN}_n${N}_e${E}_obs${OBS}_cc${CC}
  timeout 600 ncu --set full --clock-control none --cache-control $CC --import-source on -k regex:$KRE -s 4 -c 1 \
      -o gpurun_out/$NAME -f python scripts/run_cfg.py $SCEN $N $E 8 $OBS > gpurun_out/$NAME.log 2>&1
  tail -1 gpurun_out/$NAME.log
  ncu -i gpurun_out/$NAME.ncu-rep --page raw --csv > gpurun_out/$NAME.raw.csv 2>/dev/null
  ncu -i gpurun_out/$NAME.ncu-rep --page source --csv > gpurun_out/$NAME.source.csv 2>/dev/null
  gzip -f gpurun_out/$NAME.source.csv
  [ "$KEEP_REP" = "1" ] || rm -f gpurun_out/$NAME.ncu-rep
done
ls -la gpurun_out | tail -20
