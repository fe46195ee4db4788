#!/bin/bash
# A/B bench of alternative builds of the library: scripts/ab_bench.sh [steps] lib1.so lib2.so ...  ("-" = the in-tree build)
# Prints value / e2e / kernel averages per build (bench.py --no-extras --no-cpu-baseline; extra bench args in $AB_ARGS).
steps=${1:-5}; shift
mkdir -p gpurun_out
for lib in "$@"; do
  if [ "$lib" = "-" ]; then unset HALOTRACE_B200_LIB HALOTRACE_LIB; tag=intree; else export HALOTRACE_B200_LIB=$lib HALOTRACE_LIB=$lib; tag=$(basename $lib .so); fi
  python bench.py --steps $steps --warmup 3 --no-extras --no-cpu-baseline ${AB_ARGS} > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{tag}.json").read().strip().splitlines()[-1])
    pk = d["roofline"]["per_kernel"]
    print(f"{tag:24s} value {d['value']:8.1f} e2e {d['e2e']['value']:8.1f} ms/step {d['ms_per_step']:7.3f} " +
          " ".join(f"{k} {v['avg_ms']:.4f}" for k, v in pk.items()) + f" clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(tag, "FAILED", e)
PY
done
