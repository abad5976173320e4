#!/bin/bash
# Regenerates the ncu evidence bench.py quotes for K1 (run on the GPU box, under gpurun):
#   1. the bench command exits 0 without ncu,
#   2. launch list of the same command (per-launch device times: shares of the step),
#   3. one `--set full` capture of the K1 launch,
#   4. tools/ncu_k1_summary.py turns the capture into profiles/r02_k1_traffic.json
#      (DRAM bytes per launch, pipe utilisations, wavefronts) — run here or on the CPU box.
# Usage: tools/ncu_k1.sh [tag]      (outputs under gpurun_out/)
set -e
cd "$(dirname "$0")/.."
TAG=${1:-r02}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_bench_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches_cfg5.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1_resample_kernel -s 3 -c 1 \
    -o gpurun_out/${TAG}_k1 $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_k1.ncu-rep --page raw --csv > gpurun_out/${TAG}_k1_ncu_full_metrics.csv 2>/dev/null
python tools/ncu_k1_summary.py gpurun_out/${TAG}_k1_ncu_full_metrics.csv gpurun_out/${TAG}_k1_traffic.json
