#!/usr/bin/env bash
# Sweep the communicator CTA cap / SM reservation (MEDVILL_COMM_CTAS, csrc/engine.cu) on N GPUs of one box.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/sweep_comm_ctas.sh 8 "default 4 8 16"'
# One JSON line per setting lands in gpurun_out/comm_ctas_n<N>_<setting>.json; compare ms_per_step and
# roofline.gemm_ms_per_step ("default" = 8 SMs reserved, NCCL's own channel count).
set -u
N=${1:-2}
SETTINGS=${2:-"default 4 8 16"}
STEPS=${STEPS:-15}
mkdir -p gpurun_out
for c in $SETTINGS; do
  out=gpurun_out/comm_ctas_n${N}_${c}.json
  if [ "$c" = default ]; then unset MEDVILL_COMM_CTAS; else export MEDVILL_COMM_CTAS=$c; fi
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus "$N" --steps "$STEPS" --warmup 3 > "$out" 2> "${out%.json}.err"
  echo "MEDVILL_COMM_CTAS=$c rc=$? $(python - "$out" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%.1f samples/s  %.3f ms/step  gemm %.2f ms" % (d["value"], d["ms_per_step"], d["roofline"]["gemm_ms_per_step"]))
except Exception as e:
    print("no result:", e)
PY
)"
done
