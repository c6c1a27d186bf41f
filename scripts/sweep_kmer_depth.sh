python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for k in 14 13 12; do echo "K=$k"; BSQ_KMER_K=$k python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stage_ms_per_step'], d['counters_per_launch']['n_extend'])"; done
