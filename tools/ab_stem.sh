for i in 1 2; do
for v in 1 0; do
MEDVILL_STEM_GEMM=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/abe_${v}_$i.json 2> gpurun_out/abe_${v}_$i.err
python - <<PY
import json
d=json.load(open("gpurun_out/abe_${v}_$i.json")); r=d["roofline"]
print("stem_gemm=$v run $i", round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"])
PY
done; done
