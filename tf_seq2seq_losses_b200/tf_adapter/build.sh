#!/bin/bash
# Compiles the TensorFlow custom op against the installed TensorFlow (needs `import tensorflow`; not available in the
# image this repository was developed in).  Run `make -C ../csrc` first.
set -e
cd "$(dirname "$0")"
FLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags() + tf.sysconfig.get_link_flags()))')
g++ -std=c++17 -shared -fPIC ctc_b200_tf_op.cc -o ctc_b200_tf_op.so $FLAGS -I../../include -I/usr/local/cuda/include \
    -L.. -lctc_b200 -Wl,-rpath,'$ORIGIN/..'
echo "built $(pwd)/ctc_b200_tf_op.so"
