#!/bin/bash
# Round-1 evidence capture (run on the B200 box through gpurun, from the repo root).  Plain runs first; a command
# goes under ncu only after the same command has exited 0 without it.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r1_pytest_gpu.log 2>&1; echo "pytest rc=$?" 
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r1_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_reference.json 2> gpurun_out/r1_bench_reference.err; echo "reference rc=$?"
# launch list of one bench step (warm-up launches are skipped by count below when summarising)
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r1_launches_bench.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
# full captures: one GNN layer's kernels + Sinkhorn of a 40-pair forward, and SuperPoint bf16 on 4 images
python tools/prof_forward.py 40 2048 1 > gpurun_out/plain_fwd.log 2>&1 && \
  ncu --set full --clock-control none --import-source on --launch-skip 60 -c 7 -o gpurun_out/r1_superglue_kernels -f \
      python tools/prof_forward.py 40 2048 1 > gpurun_out/ncu_sg.log 2>&1; echo "sg full rc=$?"
python tools/prof_forward.py 40 2048 1 > /dev/null 2>&1 && \
  ncu --set full --clock-control none -k regex:sks_ -c 3 -o gpurun_out/r1_sinkhorn_kernels -f \
      python tools/prof_forward.py 40 2048 1 > gpurun_out/ncu_sk.log 2>&1; echo "sk full rc=$?"
python tools/prof_superpoint.py bf16 > gpurun_out/plain_sp.log 2>&1 && \
  ncu --set full --clock-control none -c 22 -o gpurun_out/r1_superpoint_kernels -f \
      python tools/prof_superpoint.py bf16 > gpurun_out/ncu_sp.log 2>&1; echo "sp full rc=$?"
tail -1 gpurun_out/r1_pytest_gpu.log; tail -1 gpurun_out/r1_smoke.log; cat gpurun_out/r1_bench.json | cut -c1-300
