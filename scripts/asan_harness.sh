#!/bin/bash
# compute-sanitizer is closed on the GPU pool; the assembly / error kernels' per-thread bodies, the
# expression compiler and the VM are host/device code, so they are run under AddressSanitizer and
# UBSan on the CPU through the test harness instead.  Run from the repository root.
set -e
cd "$(dirname "$0")/.."
B=tests/native/_build
g++ -O1 -g -std=c++17 -fPIC -shared -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off \
    -I/usr/local/cuda/include -o $B/libemul_asan.so tests/native/emul.cpp \
    dealiiscale_b200/csrc/expr_compile.cpp dealiiscale_b200/csrc/problem_programs.cpp
cp $B/libemul.so /tmp/libemul_plain.so
cp $B/libemul_asan.so $B/libemul.so
trap 'cp /tmp/libemul_plain.so '$B'/libemul.so; rm -f '$B'/libemul_asan.so' EXIT
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 \
LD_PRELOAD="$(g++ -print-file-name=libasan.so) $(g++ -print-file-name=libubsan.so)" \
    python -m pytest tests/test_emul_parity.py tests/test_expr_compile.py -x -q
