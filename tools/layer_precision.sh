#!/bin/bash
# which tensor-core pass moves the deep-gradient error of the reduced-width layer test?
for m in 0 1 2 4 7; do
  echo "== CFUN_TC_PASSES=$m"
  CFUN_TC_PASSES=$m timeout 300 python -m pytest "tests/test_gpu_model.py::test_layers_match_reference_golden[beginning]" -q -m gpu -p no:cacheprovider 2>&1 | grep -E "passed|failed|AssertionError: \(" | head -3
done
