#!/bin/bash
# GPU parity suite + one short bench line (stage times, logical extension count)
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stage_ms_per_step'], d['counters_per_launch']['n_extend'])"
