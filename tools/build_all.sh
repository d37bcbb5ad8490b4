#!/bin/bash
# Build the product and development libraries to temporary names and rename them into place (a GPU-run snapshot
# taken at any moment sees complete files).
set -e
cd "$(dirname "$0")/../dalle_mtf_b200/csrc"
make -j16 LIB=../.tmp_prod.so BUILD=build > /dev/null
make -j16 DEV=1 LIB=../.tmp_dev.so BUILD=build_dev > /dev/null
mv ../.tmp_prod.so ../libdalle_b200.so
mv ../.tmp_dev.so ../libdalle_b200_dev.so
ls -la ../*.so
