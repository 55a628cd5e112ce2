#!/usr/bin/env bash
# In-step A/B of engine switches: bench.py (cfg3 unless WORKLOAD is set) run back to back, interleaved, REPS times per variant.
# usage: bash tools/ab_step.sh "<name>:<ENV=1 ENV2=1>" ...   (name "base" with no env = default build)
mkdir -p gpurun_out
REPS=${REPS:-2}
for r in $(seq 1 $REPS); do
  for spec in "$@"; do
    name="${spec%%:*}"; envs="${spec#*:}"; [ "$envs" = "$spec" ] && envs=""
    batch=""; for kv in $envs; do case "$kv" in BATCH=*) batch="--batch ${kv#BATCH=}";; esac; done
    env $envs python bench.py --workload ${WORKLOAD:-cfg3} --steps ${STEPS:-3} --warmup ${WARMUP:-2} --no-cpu-baseline $batch > gpurun_out/ab_${name}_$r.json 2> gpurun_out/ab_${name}_$r.err
    python - "$name" "$r" <<'P'
import json, sys
n, r = sys.argv[1:3]
try:
    d = json.load(open(f"gpurun_out/ab_{n}_{r}.json"))
    rf = d["roofline"]
    print(f"{n:12s} rep {r}: {d['value']:.1f} ms  e2e {d['e2e']['value']:.1f}  sm {d['clocks']['sm_mhz']} MHz  attn {rf['avg_launch_ms']:.3f} ms/launch "
          f"(share {rf['share_of_step']:.3f})  gemm share {rf['gemm_share_of_step']:.3f} = {rf['gemm_share_of_step'] * d['value']:.0f} ms", flush=True)
except Exception as e:
    print(n, r, "failed", repr(e), flush=True)
P
  done
done
