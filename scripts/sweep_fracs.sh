#!/bin/bash
# e2e of bsq_align_batch_datums against the chunk fractions of the two-lane pipeline (BSQ_CHUNK_FRACS)
run() { python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.2f e2e %.2f ms %.2f ascii %.2f rows %.2f' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e'].get('ms_per_step', 0), d.get('e2e_ascii', {}).get('value', 0)/1e6, d.get('e2e_rows', {}).get('value', 0)/1e6))"; }
for f in "0.5,0.5" "0.4,0.6" "0.6,0.4" "0.3,0.4,0.3" "0.25,0.5,0.25" "0.2,0.6,0.2" "0.15,0.35,0.35,0.15"; do
  echo -n "fracs=$f  "
  BSQ_CHUNK_FRACS=$f run "$@"
done
