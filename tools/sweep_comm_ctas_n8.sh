#!/bin/bash
# 8-GPU sweep of the communicator CTA cap / SM reservation (MEDVILL_COMM_CTAS=n sets both; run under `gpurun --gpus 8`):
# one JSON line per setting into gpurun_out/r02_ctas_sweep.jsonl, N=1 first for the efficiency denominator.
set -u
OUT=gpurun_out/r02_ctas_sweep.jsonl
: > $OUT
run() {
  local label=$1 n=$2; shift 2
  if [ "$n" = 1 ]; then
    env "$@" python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>> gpurun_out/r02_ctas_sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); d['label']='$label'; print(json.dumps(d))" >> $OUT
  else
    env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline 2>> gpurun_out/r02_ctas_sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); d['label']='$label'; print(json.dumps(d))" >> $OUT
  fi
}
run n1 1 X=1
for c in "$@"; do run n8_ctas$c 8 MEDVILL_COMM_CTAS=$c; done
run n8_default 8 X=1
python - <<'PY'
import json
rows=[json.loads(l) for l in open("gpurun_out/r02_ctas_sweep.jsonl")]
base=rows[0]["value"]
for r in rows:
    rf=r["roofline"]
    print("%-14s n=%d value %.1f (eff %.4f) ms/step %.2f gemm %.2f e2e %.1f" % (r["label"], r["n_gpus"], r["value"], r["value"]/(base*r["n_gpus"]), r["ms_per_step"], rf["gemm_ms_per_step"], r["e2e"]["value"]))
PY
