#!/bin/bash
# Builds libmdkm of an older revision into tools/_build/libmdkm_<rev>.so (A/B timing on one box:
#   MDKM_LIB=tools/_build/libmdkm_<rev>.so python tools/sweep_opts.py CELL_PX 0 c2)
rev=$1
tmp=$(mktemp -d)
mkdir -p $tmp/pkg/csrc $tmp/include tools/_build
for f in $(git ls-tree --name-only $rev 3d-point-cloud-multiday-imagery_b200/csrc/); do git show $rev:$f > $tmp/pkg/csrc/$(basename $f); done
git show $rev:include/mdkm.h > $tmp/include/mdkm.h
(cd $tmp/pkg && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC -o libmdkm.so csrc/mdkm.cu -ldl) && cp $tmp/pkg/libmdkm.so tools/_build/libmdkm_$rev.so
rm -rf $tmp
ls -la tools/_build/libmdkm_$rev.so
