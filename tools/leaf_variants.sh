#!/bin/bash
# development only: times the leaf emitter variants (SVO_LEAF_VARIANT) on one big config
cfg=${1:-c5}
for v in 0 1 2 3 4 5; do
  echo "variant $v"; SVO_LEAF_VARIANT=$v timeout 600 python tools/profile_big.py $cfg 3 2>&1 | tail -1 | python -c "
import sys, json
l=sys.stdin.read().strip(); d=json.loads(l[l.index('{'):]); print({k: round(d[k],4) for k in ('ms_compact','ms_build','ms_emit','ms_emit_leaf')})"
done
