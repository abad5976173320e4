#!/bin/bash
# runs tools/microbench.py against alternative builds of the library (development aid)
cp scikits-bootstrap_b200/libb200boot.so /tmp/lib_base.so
echo "== base"; python tools/microbench.py 20000 2>&1 | grep resample
for f in variants_tmp/*.so; do
  cp $f scikits-bootstrap_b200/libb200boot.so
  echo "== $f"; python tools/microbench.py 20000 2>&1 | grep resample
done
cp /tmp/lib_base.so scikits-bootstrap_b200/libb200boot.so
