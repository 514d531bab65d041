#!/bin/bash
# Run on the GPU box (under gpurun): ncu captures for profiles/.  Usage: bash tools/make_profiles.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
# 1. launch list of the bench command (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
# 2. full capture of the two dominant kernels (one launch each)
ncu --set full --clock-control none --import-source on -k regex:"k_interp_tiled|k_gridding_tiled" -s 2 -c 2 \
    -o gpurun_out/prof_${TAG} python tools/prof_kernels.py 3 > gpurun_out/prof_${TAG}.log 2>&1
tail -1 gpurun_out/prof_${TAG}.log
