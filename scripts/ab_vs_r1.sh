#!/bin/bash
# Same box, same minute: the round-1 library (scripts/_build/r1_tree, built from commit 6e1e5a9) against the current one
# over the cloud sizes of real scenes.  usage: ab_vs_r1.sh <sizes> <degrees>
sizes=${1:-1.25e6,2.5e6,5e6,1e7}; degs=${2:-0 1 2 3}
export SPZB200_NO_REBUILD=1
for rep in 1 2; do
for deg in $degs; do
  (cd scripts/_build/r1_tree && python scripts/size_sweep.py $sizes $deg | sed 's/^{/{"lib": "round1", /')
  python scripts/size_sweep.py $sizes $deg | sed 's/^{/{"lib": "round2", /'
done
done
