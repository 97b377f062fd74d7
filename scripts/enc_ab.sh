#!/bin/bash
# encode-only A/B of the tile geometry at large sizes: enc_ab.sh <sizes> <degrees>
export SPZB200_NO_REBUILD=1
for rep in 1 2; do for t in 320 128; do SPZB200_TILE=$t python scripts/enc_sweep.py $1 $2 | sed "s/^{/{\"SPZB200_TILE\": \"$t\", /"; done; done
