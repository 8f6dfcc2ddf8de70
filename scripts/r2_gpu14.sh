#!/bin/bash
# K2w: parked leaves (tested together every PARK+1 steps) at 32 warps: speed, node visits per ray, parity subset
P="python scripts/profile_render.py"
export RTW_LIB=raytracing-one-weekend_b200/librtw_b200_tune.so
for rep in 1 2; do
$P --kernel bvh --spp 128 2>&1 | tail -1 | sed "s/^/park=0 /"
for pk in 1 3 7 15; do RTW_WF_PARK=$pk $P --kernel bvh --spp 128 2>&1 | tail -1 | sed "s/^/park=$pk /"; done
done
$P --kernel bvh --spp 16 --stats 2>&1 | tail -1 | sed "s/^/park=0 /"
for pk in 1 3 7 15; do RTW_WF_PARK=$pk $P --kernel bvh --spp 16 --stats 2>&1 | tail -1 | sed "s/^/park=$pk /"; done
