#!/bin/bash
# Build a tuning variant of libfmb.so with extra -D flags into tools/variants/libfmb_<name>.so
# (loaded by tools/sweep_env.py through FMB_LIB_PATH; never by the product path).
# usage: [NVXFLAGS="-Xptxas ..."] bash tools/build_variant.sh <name> [-DFOO=1 ...]   (NVXFLAGS: nvcc-only flags)
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=/tmp/fmb_variant_$NAME; rm -rf $W; mkdir -p $W $ROOT/tools/variants
cd $ROOT/rtl_fm_player_b200/csrc
for f in fm_design fm_filesrc fm_wav fm_timeshift fm_dropin fmb_multi; do gcc -O2 -std=gnu11 -ffp-contract=off -fPIC -I$ROOT/include -I. "$@" -c $f.c -o $W/$f.o; done
for f in fmb_kernels fmb_api; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC -I$ROOT/include -I. $NVXFLAGS "$@" -c $f.cu -o $W/$f.o; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/tools/variants/libfmb_$NAME.so $W/*.o -lpthread
echo built tools/variants/libfmb_$NAME.so
