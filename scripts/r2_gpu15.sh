#!/bin/bash
# K1 (sphere sweep) with the packed FFMA2 over ray pairs against the scalar build: speed for 2 and 4 rays per lane, parity tests
P="python scripts/profile_render.py"
for rep in 1 2; do
for tag in product scalar; do
  lib=raytracing-one-weekend_b200/librtw_b200_$tag.so; [ $tag = product ] && lib=""
  for rpl in 2 4; do RTW_LIB=$lib $P --kernel spheres --rays-per-lane $rpl --spp 64 2>&1 | tail -1 | sed "s/^/$tag /"; done
done; done
timeout 900 python -m pytest tests -m gpu -q -x -k "cover or sphere or sweep or determinism or depth_rule or crossover" 2>&1 | tail -3
