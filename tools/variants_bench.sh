#!/bin/bash
# Measurement helper (GPU box): bench.py for every kernel variant under build_variants/ (GVTB200_LIB_PATH hook), one line each.
for so in "" build_variants/lib_*.so; do
  name=${so:-default}
  GVTB200_LIB_PATH=${so:+$PWD/$so} timeout 150 python bench.py --no-cpu "$@" 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name', round(d['ms_per_step'], 4), [round(l['ms'], 3) for l in d['roofline']['launches']], {k: round(v['ms_per_frame'], 3) for k, v in d.get('other_workloads', {}).items()})
"
done
