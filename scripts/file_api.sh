#!/bin/bash
# The file-level C++ API (saveSpz / loadSpz in memory) and pack / unpack: this library against the reference's own sources, same box.
set -u
for n in 6e4 1e6 4e6; do scripts/_build/file_api_timing_ref $n 1 reference | grep -v "^\[SPZ"; scripts/_build/file_api_timing $n 3 spz_b200 | grep -v "^\[SPZ" | tail -2; done
scripts/_build/file_api_timing 1e7 2 spz_b200 | grep -v "^\[SPZ" | tail -1
SPZ_B200_GZIP_THREADS=1 scripts/_build/file_api_timing 1e6 2 "spz_b200 SPZ_B200_GZIP_THREADS=1" | tail -1
