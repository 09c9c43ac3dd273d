#!/bin/bash
# A/B timing of TU-path build variants on the GPU box: tools/variants_tu.sh build/libA.so build/libB.so ...
for lib in "$@"; do
  echo "== $lib"
  VVCB_LIBRARY_PATH=$PWD/$lib python tools/profile_tu.py --width 1920 --height 1080 --passes 4
done
