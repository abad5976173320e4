#!/bin/bash
# builds experimental variants of the library into variants_tmp/ (development aid)
set -e
cd "$(dirname "$0")/.."
mkdir -p variants_tmp
rm -f variants_tmp/*.so
build() { name=$1; shift; python scikits-bootstrap_b200/build.py "$@" -ovariants_tmp/$name.so > /dev/null && echo built $name; }
build unroll2 -DB2_CLASS_UNROLL=2 &
build mulhi -DB2_K1_SHF=0 &
build split -DB2_K1_PAIR=0 &
build r10 -DB2_PHILOX_ROUNDS=10 &
wait
