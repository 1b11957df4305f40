#!/bin/bash
# Development aid: build libnss_b200 with an alternative nss_eas_core.cuh into gpurun_variants/<name>.so
#   tools/build_variant.sh <name> <path-to-core-header> [<path-to-nss_eas_kernel.cu> [<path-to-nss_f64_kernels.cu>]]
#   EXTRA="-DNAME=V ..." adds flags to the EAS translation unit
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; HDR=$2
TMP=$(mktemp -d)
mkdir -p $TMP/include $TMP/nuspacesim_b200/csrc $ROOT/gpurun_variants
cp $ROOT/include/nss_b200.h $TMP/include/
cp $ROOT/nuspacesim_b200/csrc/*.cu $ROOT/nuspacesim_b200/csrc/*.h $TMP/nuspacesim_b200/csrc/
cp $HDR $TMP/nuspacesim_b200/csrc/nss_eas_core.cuh
if [ -n "$3" ]; then cp $3 $TMP/nuspacesim_b200/csrc/nss_eas_kernel.cu; fi
if [ -n "$4" ]; then cp $4 $TMP/nuspacesim_b200/csrc/nss_f64_kernels.cu; fi
cd $TMP/nuspacesim_b200/csrc
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC"
nvcc $F $EXTRA -c nss_eas_kernel.cu -o eas.o
nvcc $F -fmad=false -c nss_f64_kernels.cu -o f64.o
nvcc $F -c nss_microbench.cu -o mb.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/gpurun_variants/$NAME.so eas.o f64.o mb.o
rm -rf $TMP
echo built $ROOT/gpurun_variants/$NAME.so
