#!/bin/bash
# wall time of the C++ drop-in API on pageable std::vector planes (scripts/api_timing.cc), best-of-4 style: rep 0 includes
# context creation and the one-time pinned bounce buffers
mkdir -p scripts/_build gpurun_out
g++ -std=c++17 -O2 -pthread -Iinclude/spz scripts/api_timing.cc -o scripts/_build/api_timing -Lspz_b200/_lib -lspz_b200 -Wl,-rpath,'$ORIGIN/../../spz_b200/_lib' || exit 1
for n in 6e4 2e5 1e6 2e6 4e6 1e7; do echo "== $n gaussians SH3 (default policy)"; scripts/_build/api_timing $n 4; done
echo "== 1e7, SPZ_B200_ZEROFILL=1 (plain resize)"; SPZ_B200_ZEROFILL=1 scripts/_build/api_timing 1e7 4
echo "== 2e6, SPZB200_BOUNCE_MIN_MB=100000 (never bounce)"; SPZB200_BOUNCE_MIN_MB=100000 scripts/_build/api_timing 2e6 4
