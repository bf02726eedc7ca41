#!/bin/bash
# Round-2 evidence capture (run on the B200 box through gpurun, from the repo root).  Plain runs first; a command
# goes under ncu only after the same command has exited 0 without it.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu < /dev/null > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" < /dev/null > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 10 --warmup 5 < /dev/null > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 < /dev/null > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; echo "reference rc=$?"
# launch list of one bench step (the parity / secondary legs are switched off so that the list is the step itself)
B="bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs --no-images"
timeout 300 python $B < /dev/null > gpurun_out/plain_bench.log 2>&1 && \
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/r2f_launches_bench.csv \
      python $B < /dev/null > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
# full captures: one GNN layer's kernels + Sinkhorn of a 40-pair forward
timeout 200 python tools/prof_forward.py 40 2048 1 < /dev/null > gpurun_out/plain_fwd.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 60 -c 8 -o gpurun_out/r2f_superglue_kernels -f \
      python tools/prof_forward.py 40 2048 1 < /dev/null > gpurun_out/ncu_sg.log 2>&1; echo "sg full rc=$?"
timeout 200 python tools/prof_forward.py 40 2048 1 < /dev/null > /dev/null 2>&1 && \
  timeout 900 ncu --set full --clock-control none -k regex:sks_ -c 3 -o gpurun_out/r2f_sinkhorn_kernels -f \
      python tools/prof_forward.py 40 2048 1 < /dev/null > gpurun_out/ncu_sk.log 2>&1; echo "sk full rc=$?"
tail -1 gpurun_out/r2f_pytest_gpu.log; tail -1 gpurun_out/r2f_smoke.log; cut -c1-300 gpurun_out/r2f_bench.json
# attention kernel alone: full set with source-level stall samples (tools/ncu_stalls.py reads it)
timeout 200 python tools/prof_attention.py < /dev/null > /dev/null 2>&1 && \
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:tc_attention --launch-skip 3 -c 1 -o gpurun_out/r2f_attention -f \
      python tools/prof_attention.py < /dev/null > gpurun_out/ncu_att.log 2>&1; echo "attention full rc=$?"
