#!/bin/bash
# e2e throughput of bsq_align_batch against the chunk size of its two-lane pipeline
for c in 32768 65536 131072 262144 524288; do
  echo "chunk=$c"; BSQ_CHUNK_READS=$c python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches'])"
done
