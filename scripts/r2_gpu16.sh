#!/bin/bash
# suzanne (968 triangles, 108 KB of tables) on K2w with the tables in shared memory at 20 warps, against the per-lane kernel reading them through L1
P="python scripts/profile_render.py"
for rep in 1 2; do
$P --kernel bvh --scene suzanne --spp 64 --depth 20 2>&1 | tail -1 | sed "s/^/per-lane /"
RTW_WF_TRI_TIERS=1 $P --kernel bvh --scene suzanne --spp 64 --depth 20 2>&1 | tail -1 | sed "s/^/K2w smem /"
done
RTW_WF_TRI_TIERS=1 timeout 600 python -m pytest tests -m gpu -q -x -k "suzanne or mixed or mesh_on_ground" 2>&1 | tail -3
