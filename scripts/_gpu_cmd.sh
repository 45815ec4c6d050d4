mkdir -p gpurun_out
python bench.py > gpurun_out/r2m_bench_cfg4.json 2> gpurun_out/r2m_bench_cfg4.err; tail -2 gpurun_out/r2m_bench_cfg4.err
for w in cfg2-encoder cfg3-mae cfg3-simple-mae cfg1-vqvae; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/r2m_bench_$w.json 2> gpurun_out/r2m_bench_$w.err; tail -2 gpurun_out/r2m_bench_$w.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2m_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d.get("gpu_eager_baseline") or {}
        print(f.split("bench_")[1], round(d["value"],1), d["unit"], round(d["ms_per_step"],2),"ms e2e",round(d["e2e"]["value"],1), "eager bf16", (e.get("bf16_autocast") or {}).get("value"), "fp32", (e.get("fp32") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
