#!/bin/bash
# Profiling recipe of this repo (run under gpurun on ONE B200; see /opt/skills/guides/B200_PROFILING.md).
#   bash profiles/run_profiles.sh <tag>          e.g. r01_v4
# Produces in gpurun_out/: <tag>_launches.csv (every launch of 3 training steps with its device time) and one
# `ncu --set full` report per hot kernel.  profiles/summarize_ncu.py turns them into the committed summaries.
set -u
TAG=${1:-prof}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python profiles/prof_step.py --steps 3 > $OUT/${TAG}_launches.log 2>&1
KERNELS=${2:-"k_blend_bwd2 k_blend_fwd2 k_preprocess_bwd k_preprocess\$ k_tile_sort\$ k_scatter k_ssim_fwd k_ssim_bwd k_adam\$"}
for k in $KERNELS; do
  name=$(echo $k | tr -d '\\($')
  SKIP=2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $SKIP -c 1 -f -o $OUT/${TAG}_$name \
      python profiles/prof_step.py --steps 3 > $OUT/${TAG}_$name.log 2>&1
done
ls -la $OUT | grep $TAG
