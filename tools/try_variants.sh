#!/bin/bash
# runs tools/microbench.py against the library and every variants_tmp/*.so (development aid)
cd "$(dirname "$0")/.."
N=${1:-20000}
echo "== base"; python tools/microbench.py $N --check 2>&1 | tail -1
for f in variants_tmp/*.so; do
  echo "== $f"; B200BOOT_LIB_PATH=$PWD/$f python tools/microbench.py $N --check 2>&1 | tail -1
done
