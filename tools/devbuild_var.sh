#!/bin/bash
# usage: tools/devbuild_var.sh name "flags" [name "flags" ...]  -> csrc/libnemb200_<name>.so (S=50 development builds, in parallel)
set -e
cd "$(dirname "$0")/../nem-mcmc-optimization_b200/csrc"
FL="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
while [ $# -ge 2 ]; do
  n=$1; f=$2; shift 2
  ( nvcc $FL -DNEMB200_DEV_S50 $f -c -o /tmp/nemb200_dev/var_$n.o score_tc.cu && \
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libnemb200_$n.so /tmp/nemb200_dev/var_$n.o /tmp/nemb200_dev/{dataset,graph,mcmc,mcmc_fused,cell}.o ) &
done
wait
