#!/bin/bash
# builds one binary per configuration (in parallel) into tools/lab_bin/; usage: tools/ntt_lab.sh "R:ABL[:DEPHASE_NS]" ...
cd "$(dirname "$0")"
cfgs=${@:-4:0 4:1 4:2 4:4 4:7 4:8 4:16 4:32 4:64 4:128 4:256 4:512}
mkdir -p lab_bin
build() {
  IFS=: read r a d <<< "$1"; d=${d:-4000}
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -DLAB_R=$r -DABL=$a -DDEPHASE_NS=$d -o lab_bin/lab_${r}_${a}_${d} ntt_lab.cu
}
export -f build
echo $cfgs | tr ' ' '\n' | xargs -P 8 -I{} bash -c 'build {}'
ls lab_bin
