#!/bin/bash
# What the driver runs at round end, in one call: smoke, GPU parity suite, default bench (both arms), then the ncu launch list of the
# same bench command, one ncu --set full pass over the chain / extension / finalize kernels of a warmed step, and the C3-scale
# bench (3.1 Gbp reference, 64-bit rows).  Outputs under gpurun_out/.
set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 1500 gpurun_out/bench_reference.json
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3000 gpurun_out/bench_default.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-200
python scripts/step_launches.py gpurun_out/launches_final.csv > gpurun_out/step_final.txt; cat gpurun_out/step_final.txt
ncu --set full --clock-control none --import-source on -k regex:"ext_|sw_extend|chain_build|regs_finalize|regs_cigar" --launch-skip 75 -c 25 -o gpurun_out/stages_final -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_stages.log 2>&1
python bench.py --ref-mbp 3100 --reads 1000000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 1200 gpurun_out/bench_c3.json
