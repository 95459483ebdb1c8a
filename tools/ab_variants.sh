#!/bin/bash
# Developer tool: stage timings of kernel build variants (gpufluidsimulation_b200/lib/variants/*.so) at one size.
n=${1:-512}
for so in gpufluidsimulation_b200/lib/libbimocq_b200.so gpufluidsimulation_b200/lib/variants/*.so; do
  echo "== $so"
  BMQ_LIB=$PWD/$so python tests/perf_stage_timing.py $n 2>&1 | python -c "
import sys,json
t=sys.stdin.read()
try:
    d=json.loads(t[t.index('{'):t.rindex('}')+1]); print({k:v for k,v in d['stage_ms'].items() if k!='semilag'}, 'SUM', d['step_ms_sum'])
except Exception as e: print('ERR', t[-500:])
"
done
