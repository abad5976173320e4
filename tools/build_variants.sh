#!/bin/bash
# builds experimental variants of the library into variants_tmp/ (development aid)
set -e
cd "$(dirname "$0")/.."
mkdir -p variants_tmp
rm -f variants_tmp/*.so
build() { name=$1; shift; python scikits-bootstrap_b200/build.py "$@" -ovariants_tmp/$name.so > /dev/null && echo built $name; }
build pair -DB2_K1_PAIR=1 &
build shf -DB2_K1_SHF=1 &
build pair_shf -DB2_K1_PAIR=1 -DB2_K1_SHF=1 &
build pair_pipe -DB2_K1_PIPE=1 -DB2_K1_PAIR=1 &
wait
build tree256 -DB2_TREE_BATCH=256 &
build r7 -DB2_PHILOX_ROUNDS=7 &
build r7_pair -DB2_PHILOX_ROUNDS=7 -DB2_K1_PAIR=1 &
build r7_pair_shf -DB2_PHILOX_ROUNDS=7 -DB2_K1_PAIR=1 -DB2_K1_SHF=1 &
wait
