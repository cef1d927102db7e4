#!/bin/bash
# build a named variant of libzsv_b200.so for A/B runs on one box (ZSV_LIB_PATH=...): tools/build_variant.sh NAME [git-rev]
# sources: the working tree, or `git-rev` when given.  Output: zeroshotvideoclassification_b200/build/variants/NAME.so
set -e
name=$1; rev=$2; extra=$3   # extra: additional nvcc flags, e.g. -DZSV_NO_SEG32 (use '' for rev to keep the working tree)
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/zeroshotvideoclassification_b200/build/variants
mkdir -p $out /tmp/zsv_variant_$name/csrc /tmp/zsv_variant_$name/include
if [ -n "$rev" ]; then
  for f in $(git -C $root ls-tree --name-only $rev zeroshotvideoclassification_b200/csrc/); do git -C $root show $rev:$f > /tmp/zsv_variant_$name/csrc/$(basename $f); done
  git -C $root show $rev:include/zsv_b200.h > /tmp/zsv_variant_$name/include/zsv_b200.h
else
  cp $root/zeroshotvideoclassification_b200/csrc/* /tmp/zsv_variant_$name/csrc/; cp $root/include/zsv_b200.h /tmp/zsv_variant_$name/include/
fi
cd /tmp/zsv_variant_$name/csrc
sed -i 's#"../../include/zsv_b200.h"#"../include/zsv_b200.h"#' *.h *.cu *.cuh 2>/dev/null || true
objs=""
for src in zsv_common.cu zsv_conv.cu zsv_elementwise.cu zsv_head.cu zsv_linear.cu zsv_optim.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -cudart static -I/tmp/zsv_variant_$name/include $extra -c $src -o $src.o &
  objs="$objs $src.o"
done
wait
nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -o $out/$name.so $objs
echo $out/$name.so
