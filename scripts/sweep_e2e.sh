#!/bin/bash
# e2e of bsq_align_batch_datums against chunk size, lane stream priorities (BSQ_NO_STREAM_PRIO) and chunk ordering (BSQ_CHUNK_SERIAL)
run() { python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.2f e2e %.2f ms %.2f ascii %.2f rows %.2f' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e'].get('ms_per_step', 0), d.get('e2e_ascii', {}).get('value', 0)/1e6, d.get('e2e_rows', {}).get('value', 0)/1e6))"; }
for c in 0 349526 262144; do
  for p in 1 0; do
    echo -n "chunk=$c prio=$p  "
    if [ $p = 0 ]; then export BSQ_NO_STREAM_PRIO=1; else unset BSQ_NO_STREAM_PRIO; fi
    if [ $c = 0 ]; then unset BSQ_CHUNK_READS; else export BSQ_CHUNK_READS=$c; fi
    run "$@"
  done
done
