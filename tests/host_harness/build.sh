#!/bin/bash
# usage: build.sh <out.so> [extra -D flags]   (TEST INFRASTRUCTURE: host build of the kernels' per-pair arithmetic)
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$1; shift
g++ -O2 -std=c++17 -ffp-contract=off -mfma -fPIC -shared -w "$@" -I/usr/local/cuda/include -I$HERE/../../lasr_b200/csrc -I$HERE/../../include -o $OUT $HERE/raster_math_host.cpp
