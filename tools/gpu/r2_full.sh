# round 2: the full single-GPU validation: pytest -m gpu, smoke, default bench + reference arm, the other BASELINE
# configurations with roofline + parity, ncu launch list and full captures -> gpurun_out/ (copied to profiles/r02/)
mkdir -p gpurun_out/r02
O=gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total,driver_version --format=csv > $O/gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -q -n 4 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -n 6 $O/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke ) > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 $O/smoke.log
( time timeout 900 python bench.py ) > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -n 3 $O/bench.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; echo "ref rc=$?"
b() { name=$1; shift; timeout 900 python bench.py "$@" > $O/bench_$name.json 2> $O/bench_$name.err; echo "$name rc=$?"; }
b K0 --camera K0 --cpu-row-stride 4
b K1 --camera K1 --cpu-row-stride 4
b alpha1.0 --alpha 1.0 --cpu-row-stride 4
b nearest --filter nearest --cpu-row-stride 4
b mip --mip --alpha 0.5 --cpu-row-stride 8
b C1 --config C1
b C2 --config C2
b C2_window --config C2 --window 30 180
b C3 --config C3 --alpha 0.05 --cpu-row-stride 4
b C3_skipoff --config C3 --alpha 0.05 --skip off --no-cpu-baseline
b C3_tf --config C3 --alpha 0.05 --tf --cpu-row-stride 4
b C3_tf_skipoff --config C3 --alpha 0.05 --tf --skip off --no-cpu-baseline
b C3_nearest --config C3 --alpha 0.05 --filter nearest --cpu-row-stride 4
b C4_window --window 1000 3000 --alpha 0.05 --cpu-row-stride 4
b C4_window_skipoff --window 1000 3000 --alpha 0.05 --skip off --no-cpu-baseline
b C5 --config C5 --cpu-row-stride 8 --steps 10
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/r02/bench_*.json")) + ["gpurun_out/r02/bench.json"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(os.path.basename(f), "unreadable", e); continue
    r = d.get("roofline") or {}; c = d.get("cpu_baseline") or {}; de = d.get("dense") or {}
    print("%-28s value %8.1f  kernel %7.3f ms (%s skip=%s) frac %s  e2e %8.1f  dense %s ms  cpu %s (%s) exact=%s" % (
        os.path.basename(f), d["value"], r.get("kernel_ms_avg", 0), r.get("kernel"), r.get("empty_space_skipping"),
        ("%.4f" % r["frac"]) if r.get("frac") else None, d["e2e"]["value"], ("%.3f" % de["kernel_ms_avg"]) if de else None,
        ("%.3f" % c["value"]) if c else None, c.get("kind"), c.get("parity_bit_exact_on_sample")))
PY
