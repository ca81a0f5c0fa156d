#!/bin/bash
# A/B two builds of the library on the SAME box (box-to-box clock differences under the power cap are ~4 %, larger than most
# kernel changes): build the two variants here as libmedvill_new.so / libmedvill_old.so next to libmedvill_sm100.so, then
#   gpurun -- 'bash tools/ab_lib.sh'
# runs bench.py twice per variant, interleaved, and prints one line each.
P=multi-modality-self-supervision_b200
for i in 1 2; do
  for v in new old; do
    cp $P/libmedvill_$v.so $P/libmedvill_sm100.so
    python bench.py --steps 10 --warmup 3 > gpurun_out/ab_${v}_$i.json 2> gpurun_out/ab_${v}_$i.err
    python - <<PY
import json
d = json.load(open("gpurun_out/ab_${v}_$i.json")); r = d["roofline"]
print("$v $i", round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "gemm", round(r["gemm_ms_per_step"], 3),
      "attn", round(r["attn_fwd_ms_per_step"], 3), round(r["attn_bwd_ms_per_step"], 3), d["clocks"]["sm_mhz"])
PY
  done
done
