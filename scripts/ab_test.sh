#!/bin/bash
# usage: scripts/ab_test.sh name — run the GPU parity tests against ab/<name>.so (tuning aid; bounded by timeout)
RTNW_LIB=ab/$1.so timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
