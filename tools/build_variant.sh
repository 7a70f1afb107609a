#!/bin/bash
# Tuning builds: tools/build_variant.sh NAME "-DGMB_MIN_BLOCKS5=3 ..." -> genmap_b200/lib/variants/libgenmap_b200_NAME.so
# (kernel translation units rebuilt with the extra flags, the rest linked from build/obj).  Select at run time with
# GMB_LIB_PATH=...; never used by the product path.
set -e
NAME=$1; FLAGS=$2
cd "$(dirname "$0")/.."
python -m genmap_b200._build > /dev/null
mkdir -p build/variants/$NAME genmap_b200/lib/variants
COMMON="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Wno-deprecated-declarations"
for f in map_kernel locate_kernel exact_kernel block_kernel; do
  nvcc $COMMON $FLAGS -c -o build/variants/$NAME/${f}_cu.o genmap_b200/csrc/$f.cu &
done
wait
nvcc $COMMON -shared -o genmap_b200/lib/variants/libgenmap_b200_$NAME.so build/variants/$NAME/map_kernel_cu.o build/variants/$NAME/locate_kernel_cu.o build/variants/$NAME/exact_kernel_cu.o build/variants/$NAME/block_kernel_cu.o \
  build/obj/capi_cu.o build/obj/jump_table_cu.o build/obj/index_build_gpu_cu.o build/obj/gmb_host_cpp.o build/obj/rle_kernel_cu.o build/obj/seqan_export_cpp.o
echo genmap_b200/lib/variants/libgenmap_b200_$NAME.so
