#!/bin/bash
# 8-GPU sweep of the gradient exchange settings (run under `gpurun --gpus 8`): one JSON line per configuration into
# gpurun_out/r02_scale_sweep.jsonl.  N=1 first (same box) for the efficiency denominator.
set -u
OUT=gpurun_out/r02_scale_sweep.jsonl
: > $OUT
run() {  # label, nproc, env...
  local label=$1 n=$2; shift 2
  if [ "$n" = 1 ]; then
    env "$@" python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>> gpurun_out/r02_scale_sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); d['label']='$label'; print(json.dumps(d))" >> $OUT
  else
    env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline 2>> gpurun_out/r02_scale_sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); d['label']='$label'; print(json.dumps(d))" >> $OUT
  fi
}
run n1 1 X=1
run n8_fp32_default 8 X=1
run n8_bf16 8 MEDVILL_GRAD_COMM=bf16
run n8_bf16_ctas4 8 MEDVILL_GRAD_COMM=bf16 MEDVILL_COMM_CTAS=4
run n8_fp32_ctas4 8 MEDVILL_COMM_CTAS=4
python - <<'PY'
import json
rows=[json.loads(l) for l in open("gpurun_out/r02_scale_sweep.jsonl")]
base=rows[0]["value"]
for r in rows:
    rf=r["roofline"]
    print("%-18s n=%d value %.1f (eff %.3f) ms/step %.2f gemm %.2f e2e %.1f" % (r["label"], r["n_gpus"], r["value"], r["value"]/(base*r["n_gpus"]), r["ms_per_step"], rf["gemm_ms_per_step"], r["e2e"]["value"]))
PY
