#!/bin/bash
# Developer-only: build the kernels for HOST emulation (no GPU needed).  See README.md.
set -e
cd "$(dirname "$0")/../.."
mkdir -p tools/emu/_build
SRCS=$(ls mannequin2_b200/csrc/*.cu)
g++ -std=c++20 -O1 -g -pthread -fPIC -shared -ffp-contract=off -DMQ_EMU -Itools/emu \
    -Wno-unknown-pragmas -Wno-attributes \
    $(for s in $SRCS; do echo "-x c++ $s"; done) -o tools/emu/_build/libmq_emu.so
echo tools/emu/_build/libmq_emu.so
