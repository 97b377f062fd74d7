#!/bin/bash
# wall time of the C++ drop-in API on pageable std::vector planes (scripts/api_timing.cc)
mkdir -p scripts/_build gpurun_out
g++ -std=c++17 -O2 -pthread -Iinclude/spz scripts/api_timing.cc -o scripts/_build/api_timing -Lspz_b200/_lib -lspz_b200 -Wl,-rpath,'$ORIGIN/../../spz_b200/_lib' || exit 1
scripts/_build/api_timing 1e7 5
echo "== SPZB200_BOUNCE_MIN_MB=0 (always bounce)"; SPZB200_BOUNCE_MIN_MB=0 scripts/_build/api_timing 1e7 5
echo "== 2M points"; scripts/_build/api_timing 2e6 5
echo "== SPZ_B200_ZEROFILL=1 (plain resize)"; SPZ_B200_ZEROFILL=1 scripts/_build/api_timing 1e7 5
