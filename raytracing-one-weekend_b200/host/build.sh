#!/bin/bash
# Builds the host side in tree: librtweekend_host.so (scene model + render() + C hooks) and the `rtweekend` executable, both with the
# default (virtual) primitive model, and the same pair with -DRTWEEKEND_USE_VARIANT_PRIMITIVES (the std::variant model the reference
# selects with that macro, primitive-model.h:1-2): librtweekend_host_variant.so, rtweekend_variant.
set -euo pipefail
cd "$(dirname "$0")"
CXX=/usr/bin/g++
[ -x "$CXX" ] || CXX=g++
FLAGS="-std=c++20 -O2 -fPIC -Wall -Wextra -pedantic -fvisibility=hidden"
LIB=../librtweekend_host.so
EXE=../rtweekend
LIBV=../librtweekend_host_variant.so
EXEV=../rtweekend_variant
SRC="render.cpp common-model.cpp random-utils.cpp scenes.cpp obj-loader.cpp"
newer=0
for f in $SRC capi.cpp main.cpp *.h ../../include/rtw_b200.h build.sh ../librtw_b200.so; do
  for o in $LIB $EXE $LIBV $EXEV; do
    if [ ! -e "$o" ] || [ "$f" -nt "$o" ]; then newer=1; fi
  done
done
if [ "$newer" = 0 ] && [ "${1:-}" != "-f" ]; then echo "host up to date"; exit 0; fi
$CXX $FLAGS -shared -o $LIB $SRC capi.cpp -L.. -lrtw_b200 -Wl,-rpath,'$ORIGIN' & p1=$!
$CXX $FLAGS -o $EXE main.cpp $SRC -L.. -lrtw_b200 -Wl,-rpath,'$ORIGIN' & p2=$!
$CXX $FLAGS -DRTWEEKEND_USE_VARIANT_PRIMITIVES -shared -o $LIBV $SRC capi.cpp -L.. -lrtw_b200 -Wl,-rpath,'$ORIGIN' & p3=$!
$CXX $FLAGS -DRTWEEKEND_USE_VARIANT_PRIMITIVES -o $EXEV main.cpp $SRC -L.. -lrtw_b200 -Wl,-rpath,'$ORIGIN' & p4=$!
wait $p1; wait $p2; wait $p3; wait $p4
echo "built $LIB $EXE $LIBV $EXEV"
