mkdir -p gpurun_out
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "2" ]; then timeout 600 $TR scripts/gpu_ddp_check.py > gpurun_out/r2n_ddp_check.log 2>&1; tail -6 gpurun_out/r2n_ddp_check.log; fi
timeout 900 $TR bench.py --gpus $N --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2n_bench_n$N.json 2> gpurun_out/r2n_bench_n$N.err; tail -2 gpurun_out/r2n_bench_n$N.err
timeout 900 $TR bench.py --gpus $N --steps 8 --warmup 3 --scaling strong --no-extras --no-cpu-baseline > gpurun_out/r2n_bench_strong_n$N.json 2> gpurun_out/r2n_bench_strong_n$N.err; tail -2 gpurun_out/r2n_bench_strong_n$N.err
python - <<PY
import json
for f in ["gpurun_out/r2n_bench_n$N.json","gpurun_out/r2n_bench_strong_n$N.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["scaling"], round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), d["config"]["trials_per_gpu"], d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
