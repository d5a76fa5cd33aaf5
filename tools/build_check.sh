#!/usr/bin/env bash
# Local gate before any gpurun: rebuild the product library from scratch and run the CPU suite.
set -e
cd "$(dirname "$0")/.."
make -s -C tc-resnet_b200/csrc clean
make -s -C tc-resnet_b200/csrc 2>&1 | grep -E "error" && { echo "BUILD FAILED"; exit 1; } || true
test -f tc-resnet_b200/libtcr_b200.so || { echo "BUILD FAILED (no .so)"; exit 1; }
python -m pytest tests -x -q -m "not gpu" > /tmp/build_check_pytest.txt 2>&1 || { tail -15 /tmp/build_check_pytest.txt; echo "CPU TESTS FAILED"; exit 1; }
tail -1 /tmp/build_check_pytest.txt
echo "BUILD+CPU TESTS OK"
