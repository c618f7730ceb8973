#!/bin/bash
# sweep the batch pipeline shape (streams x frames per chunk) -- experiment helper
for cfg in "8 8" "4 16" "8 4" "8 2" "4 8" "2 32" "1 64"; do
  set -- $cfg
  ORBX_PIPE=$1 ORBX_CHUNK=$2 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-matchers 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pipe $1 chunk $2: value %.0f  ms/step %.3f  e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done
