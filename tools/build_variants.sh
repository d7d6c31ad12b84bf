#!/bin/bash
# usage: tools/build_variants.sh name1 "FLAGS1" name2 "FLAGS2" ...   -> vbuild/<name>.so (travels with gpurun; git-ignored)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $ROOT/vbuild
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  ( d=$(mktemp -d); cp -r $ROOT/raytracer-utah_b200 $d/pkg; cp -r $ROOT/include $d/include; rm -rf $d/pkg/build $d/pkg/librtu_b200.so
    make -C $d/pkg -j4 EXTRA="$flags" $d/pkg/librtu_b200.so > $d/log 2>&1 || { tail -20 $d/log; exit 1; }
    cp $d/pkg/librtu_b200.so $ROOT/vbuild/$name.so; grep -h "registers\|spill" $d/pkg/build/rtu_kernels.ptxas.log | paste - - | grep -A0 "k_extend\|k_shadow_wave" | head -0; rm -rf $d; echo built $name ) &
done
wait
