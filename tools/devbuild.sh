#!/bin/bash
# Development build: csrc/libnemb200_dev.so with only the S=50 instantiations of the score kernel (30 s instead of 90 s).
# usage: tools/devbuild.sh [extra nvcc flags]; run with NEMB200_LIB_VARIANT=dev
set -e
cd "$(dirname "$0")/../nem-mcmc-optimization_b200/csrc"
mkdir -p /tmp/nemb200_dev
FL="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
for f in dataset graph mcmc mcmc_fused cell; do
  if [ ! -f /tmp/nemb200_dev/$f.o ] || [ $f.cu -nt /tmp/nemb200_dev/$f.o ] || [ common.cuh -nt /tmp/nemb200_dev/$f.o ]; then nvcc $FL -c -o /tmp/nemb200_dev/$f.o $f.cu & fi
done
nvcc $FL -DNEMB200_DEV_S50 "$@" -c -o /tmp/nemb200_dev/score_tc.o score_tc.cu
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libnemb200_dev.so /tmp/nemb200_dev/{score_tc,dataset,graph,mcmc,mcmc_fused,cell}.o
