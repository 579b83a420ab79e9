# round 2, run 59 (8 GPUs): the driver's scaling sequence on the final tree - bench.py at N = 1, 2, 4, 8 with its launch line
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_run59_bench_cfg2_1gpu.json 2> gpurun_out/r02_run59_bench_1gpu.err; echo "N=1 rc=$?"
for n in 2 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_run59_bench_cfg2_${n}gpu.json 2> gpurun_out/r02_run59_bench_${n}gpu.err; echo "N=$n rc=$?"
done
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    try:
        txt = open(f"gpurun_out/r02_run59_bench_cfg2_{n}gpu.json").read(); d = json.loads([l for l in txt.splitlines() if l.startswith("{")][0])
        print(f"N={n}", d["scaling"], round(d["value"], 1), "audio-s/s", round(d["ms_per_step"], 3), "ms/step; e2e", round(d["e2e"]["value"], 1), "; weak", (d.get("weak") or {}).get("value"), d["clocks"])
    except Exception as e:
        print(f"N={n} failed", e)
PY
