#!/bin/sh
# A/B builds of one kernel file: tools/build_ab.sh NAME path/to/variant_of_probe_fast2.cu [object name, default probe_fast2.cu.o]
# -> bbtools_b200/csrc/build/ab/libNAME.so (all other objects from the current build). Select with BBDUK_B200_LIB=...
set -e
cd "$(dirname "$0")/../bbtools_b200/csrc"
NAME=$1; SRC=$2; OBJ=${3:-probe_fast2.cu.o}
mkdir -p build/ab
cp "$SRC" build/ab/_$NAME.cu
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -I. -c build/ab/_$NAME.cu -o build/ab/$NAME.o
OTHERS=$(ls build/*.o | grep -v "/$OBJ")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/ab/lib$NAME.so build/ab/$NAME.o $OTHERS -lcudart_static -lpthread -ldl -lrt
rm -f build/ab/_$NAME.cu build/ab/$NAME.o
echo built build/ab/lib$NAME.so
