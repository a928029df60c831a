#!/bin/bash
# kernel time of bench.py's dominant kernel for each experimental library variant (tools/build_variant.sh)
cd "$(dirname "$0")/.."
for v in "$@"; do
  lib=""; [ "$v" != "base" ] && lib="$PWD/nfempy_b200/csrc/variants/libnfa_$v.so"
  NFA_LIB=${lib:-$PWD/nfempy_b200/csrc/libnfa.so} python bench.py --no-e2e --no-cpu --steps 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'step_ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4))"
done
