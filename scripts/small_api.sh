#!/bin/bash
# Development A/B: the C++ API on small pageable clouds, default staging policy against always-bounce (SPZB200_BOUNCE_MIN_MB=0).
set -u
export SPZ_B200_UNPACK_READAHEAD=0   # keeps api_timing's walk section short
for n in 2e4 6e4 1e5; do
  echo "== $n default"; scripts/_build/api_timing $n 6 | grep packGaussians | tail -3
  echo "== $n SPZB200_BOUNCE_MIN_MB=0"; SPZB200_BOUNCE_MIN_MB=0 scripts/_build/api_timing $n 6 | grep packGaussians | tail -3
done
