#!/bin/bash
# Kernel experiments: build one library per nvcc flag set into exp_libs/ (here), then on the GPU box
# `bash tools/variants.sh run` swaps each in and prints the debug accounting + one-way throughput.
#   bash tools/variants.sh build name1 "-DFOO=1" name2 "-DFOO=2" ...
cd "$(dirname "$0")/.."
if [ "$1" = build ]; then
    shift; rm -rf exp_libs; mkdir -p exp_libs
    while [ $# -ge 2 ]; do
        MSFM_NVCC_EXTRA="$2" python -m metricsfm_b200.build --force 2>&1 | grep -E " error" ; cp metricsfm_b200/csrc/libmsfm_match.so exp_libs/$1.so; shift 2
    done
    python -m metricsfm_b200.build --force > /dev/null 2>&1
else
    cp metricsfm_b200/csrc/libmsfm_match.so /tmp/orig.so
    for f in exp_libs/*.so; do
        echo "=== $(basename $f .so)"; cp $f metricsfm_b200/csrc/libmsfm_match.so; bash tools/dbg8.sh
    done
    cp /tmp/orig.so metricsfm_b200/csrc/libmsfm_match.so
fi
