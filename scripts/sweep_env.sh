#!/bin/bash
# usage: sweep_env.sh VAR v1 v2 ... : one short c2 bench per value, prints value, e2e and stage times
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$var=$v', round(d['value']/1e6,2), round(d['e2e']['value']/1e6,2), {k: round(x,2) for k,x in d['stage_ms_per_step'].items()})"
done
