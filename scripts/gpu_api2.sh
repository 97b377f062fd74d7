#!/bin/bash
g++ -std=c++17 -O2 -pthread -Iinclude/spz scripts/api_timing.cc -o scripts/_build/api_timing -Lspz_b200/_lib -lspz_b200 -Wl,-rpath,'$ORIGIN/../../spz_b200/_lib' || exit 1
for n in 2e5 1e6 2e6 4e6; do
echo "== $n direct"; SPZB200_BOUNCE_MIN_MB=100000 scripts/_build/api_timing $n 4
echo "== $n bounce"; SPZB200_BOUNCE_MIN_MB=0 scripts/_build/api_timing $n 4
done
