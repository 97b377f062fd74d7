#!/bin/bash
# encode-only timing of the shipped library and every variant under scripts/_build/variants: enc_variants.sh <sizes> <degrees> [extra env assignment]
export SPZB200_NO_REBUILD=1
[ -n "$3" ] && export "$3"
for rep in 1 2; do
  python scripts/enc_sweep.py $1 $2 | sed 's/^{/{"variant": "shipped", /'
  for lib in scripts/_build/variants/libspz_*.so; do
    v=$(basename $lib .so); SPZB200_LIB=$lib python scripts/enc_sweep.py $1 $2 | sed "s/^{/{\"variant\": \"${v#libspz_}\", /"
  done
done
