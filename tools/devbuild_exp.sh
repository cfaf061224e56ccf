#!/bin/bash
# builds csrc/libnemb200_expN.so for each N given (development skip-experiments, -DNEMB200_EXP=N), in parallel
set -e
cd "$(dirname "$0")/../nem-mcmc-optimization_b200/csrc"
FL="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
for n in "$@"; do
  ( nvcc $FL -DNEMB200_DEV_S50 -DNEMB200_EXP=$n -c -o /tmp/nemb200_dev/score_tc_exp$n.o score_tc.cu && \
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libnemb200_exp$n.so /tmp/nemb200_dev/score_tc_exp$n.o /tmp/nemb200_dev/{dataset,graph,mcmc,mcmc_fused,cell}.o ) &
done
wait
