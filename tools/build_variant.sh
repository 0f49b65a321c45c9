#!/bin/bash
# Developer tool: build an A/B variant of the library with extra -D flags applied to the fused-kernel translation units.
#   tools/build_variant.sh NAME "-DFLAG=1 ..." [files...]   ->  tf_seq2seq_losses_b200/libctc_b200_NAME.so
# (run `make` in csrc first; every translation unit not listed is taken from the regular build)
set -e
cd "$(dirname "$0")/../tf_seq2seq_losses_b200/csrc"
NAME=$1; FLAGS=$2; shift 2
FILES=${@:-kf_fused_simple_tma.cu kf_fused_classic_tma.cu}
mkdir -p /tmp/variant_$NAME
OBJS=""
for f in *.cu; do
  if echo " $FILES " | grep -q " $f "; then
    nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $FLAGS -c $f -o /tmp/variant_$NAME/${f%.cu}.o &
    OBJS="$OBJS /tmp/variant_$NAME/${f%.cu}.o"
  else
    OBJS="$OBJS ${f%.cu}.o"
  fi
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libctc_b200_$NAME.so $OBJS -lcudart
echo built ../libctc_b200_$NAME.so
