#!/bin/bash
# e2e of bsq_align_batch_datums against chunk size and chunk ordering (BSQ_CHUNK_SERIAL)
run() { python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.2f e2e %.2f ms %.2f ascii %.2f' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e'].get('ms_per_step', 0), d.get('e2e_ascii', {}).get('value', 0)/1e6))"; }
for c in 0 524288 349526 262144 200000; do
  for s in 0 1; do
    echo -n "chunk=$c serial=$s  "
    if [ $s = 1 ]; then export BSQ_CHUNK_SERIAL=1; else unset BSQ_CHUNK_SERIAL; fi
    if [ $c = 0 ]; then unset BSQ_CHUNK_READS; else export BSQ_CHUNK_READS=$c; fi
    run "$@"
  done
done
