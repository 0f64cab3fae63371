#!/bin/bash
# Runs the reference's own unit test binary (oracle/_ref/unittest64: test/main.cpp of the reference, linked against libhiprt_b200.so)
# on the GPU box. $HIPRT_PATH is what that binary reads its kernel sources, headers and golden pictures from: the reference's test/
# data staged by `make -C oracle/ref unittest`, and THIS repo's public headers as hiprt/.
# usage: tools/run_reference_unittest.sh <log file> [gtest filter]
set -u
root=$(cd "$(dirname "$0")/.." && pwd)
ut=$root/oracle/_ref/unittest_root
rm -rf "$ut/hiprt" && cp -r "$root/include/hiprt" "$ut/hiprt"
work=$(mktemp -d)
cd "$work"
HIPRT_PATH=$ut timeout ${UT_TIMEOUT:-900} "$root/oracle/_ref/unittest64" --gtest_filter="${2:--*PerformanceTest*}" > "$1" 2>&1
echo "exit code $?" >> "$1"
