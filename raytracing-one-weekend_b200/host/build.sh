#!/bin/bash
# Builds the host side in tree: librtweekend_host.so (scene model + render() + C hooks) and the `rtweekend` executable.
set -euo pipefail
cd "$(dirname "$0")"
CXX=/usr/bin/g++
[ -x "$CXX" ] || CXX=g++
FLAGS="-std=c++20 -O2 -fPIC -Wall -Wextra -pedantic -fvisibility=hidden"
LIB=../librtweekend_host.so
EXE=../rtweekend
SRC="render.cpp common-model.cpp random-utils.cpp scenes.cpp obj-loader.cpp"
newer=0
for f in $SRC capi.cpp main.cpp *.h ../../include/rtw_b200.h build.sh ../librtw_b200.so; do
  if [ ! -e "$LIB" ] || [ ! -e "$EXE" ] || [ "$f" -nt "$LIB" ] || [ "$f" -nt "$EXE" ]; then newer=1; fi
done
if [ "$newer" = 0 ] && [ "${1:-}" != "-f" ]; then echo "host up to date"; exit 0; fi
$CXX $FLAGS -shared -o $LIB $SRC capi.cpp -L.. -lrtw_b200 -Wl,-rpath,'$ORIGIN'
$CXX $FLAGS -o $EXE main.cpp $SRC -L.. -lrtw_b200 -Wl,-rpath,'$ORIGIN'
echo "built $LIB $EXE"
