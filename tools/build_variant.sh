#!/bin/bash
# build a variant of libbbx.so with extra compile flags for A/B runs: tools/build_variant.sh <name> <flags...>
# -> bbcat-dsp_b200/variants/libbbx_<name>.so (load it with BBX_LIB=...)
name=$1; shift
cd "$(dirname "$0")/../bbcat-dsp_b200/csrc"
mkdir -p ../variants
make OUT=../variants/libbbx_$name.so OBJDIR=../../build/obj_$name \
  NVFLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v $*" 2>&1 | grep -E "error|rror:" 
grep -A3 "k_irfft8ILi512\|k_rfft8ILi512\|k_fdl_mac_tbw" ../../build/obj_$name/engine.log | grep -E "registers|spill" 
