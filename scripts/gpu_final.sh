#!/bin/bash
# Round-end evidence run on one B200 (under gpurun): full GPU test suite, the default bench line with its
# sub-records, the reference arm, the ncu launch list of one bench step and one --set full capture of the
# generator + three bounce launches. Outputs land in gpurun_out/ (copied to profiles/ by hand).
tag=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_gputests.log 2>&1; tail -3 gpurun_out/${tag}_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log | cut -c1-200
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_config2.json 2> gpurun_out/${tag}_bench_config2.err; cut -c1-300 gpurun_out/${tag}_bench_config2.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err; cut -c1-300 gpurun_out/${tag}_bench_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${tag}_launches_bench_step.csv \
  python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gen_kernel|bounce_kernel" -c 4 -o gpurun_out/${tag}_c2 -f \
  python scripts/profile_step.py 16777216 1 > gpurun_out/${tag}_c2.log 2>&1; tail -1 gpurun_out/${tag}_c2.log
# grid-size check (CTAs per SM of the trace kernels; default = 3 waves of the co-resident CTAs = 12)
for b in 8 16; do AB_ARGS="--set blocks_per_sm=$b" bash scripts/ab_bench.sh 5 - | sed "s/^intree/blocks_per_sm=$b/"; done
for w in config5 pyramid; do AB_ARGS="--workload $w" bash scripts/ab_bench.sh 2 - | sed "s/^intree/$w/"; done
