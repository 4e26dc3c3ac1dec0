#!/usr/bin/env bash
# quick ncu pass: scheduler / warp-state / compute / occupancy / memory sections of the edge-search kernel only
READS=${1:-10000000}
TAG=${2:-quick}
mkdir -p gpurun_out
timeout 900 ncu --section SchedulerStats --section WarpStateStats --section ComputeWorkloadAnalysis --section Occupancy --section MemoryWorkloadAnalysis --section LaunchStats --section SpeedOfLight --section SourceCounters \
    --clock-control none --import-source on -k regex:"k_edges_probe|k_edges_verify" -s 2 -c 2 -f -o gpurun_out/${TAG}_prof python profiles/profile_step.py $READS 1 > gpurun_out/${TAG}_full.log 2>&1 < /dev/null
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page details 2>/dev/null | grep -v "^ *$" | head -250 > gpurun_out/${TAG}_details.txt
