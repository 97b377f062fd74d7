#!/bin/bash
# Development A/B: spz::packGaussians / unpackGaussians on pageable std::vector planes; copy threads and range size of the bounced pipeline.
set -u
export SPZ_B200_UNPACK_READAHEAD=0   # keeps api_timing's walk section short
run() { echo "== $*"; env "$@" | grep packGaussians | tail -2; }
for n in 6e4 2e5 1e6 4e6 1e7; do run scripts/_build/api_timing $n 5; done
for t in 4 12 16; do run SPZB200_COPY_THREADS=$t scripts/_build/api_timing 1e7 4; done
for c in 131072 524288 1048576; do run SPZB200_PAGEABLE_CHUNK_POINTS=$c scripts/_build/api_timing 1e7 4; done
run SPZB200_COPY_THREADS=12 SPZB200_PAGEABLE_CHUNK_POINTS=524288 scripts/_build/api_timing 1e7 4
run SPZB200_NT_COPY=0 scripts/_build/api_timing 1e7 4
