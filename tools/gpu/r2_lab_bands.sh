mkdir -p gpurun_out
for cfg in "4 0" "4 1" "8 0" "8 1" "2 0" "6 0"; do
  set -- $cfg
  if [ "$2" = "1" ]; then export VR_BAND_NOPRIO=1; else unset VR_BAND_NOPRIO; fi
  VR_BANDS=$1 timeout 300 python bench.py --no-cpu-baseline --no-count --no-dense --steps 30 > gpurun_out/bands_$1_$2.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/bands_$1_$2.json").read().strip().splitlines()[-1])
print("bands $1 noprio $2: e2e %.1f Mrays/s (%.3f ms)  device %.1f (%.3f ms)" % (d["e2e"]["value"], d["e2e"]["ms_per_step"], d["value"], d["ms_per_step"]))
PY
done
