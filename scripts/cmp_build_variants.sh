#!/bin/bash
# Dev tool: build times (A alone, B alone, both) of every library variant
for f in solidboolean_b200/lib/libsolidboolean_b200.so solidboolean_b200/lib/variants/libsb_*.so; do echo $f
SB_LIB_PATH=$PWD/$f python scripts/build_times.py c3
done
