# round 2, run 20: full GPU suite on the tree with 64-k stages (PS_PAIR_SUB), the cached signature walk without the reference
# cycle and the collector held off during graph capture; smoke; default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_run20_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_run20_pytest.log; tail -6 gpurun_out/r02_run20_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_run20_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_run20_smoke.log
timeout 600 python bench.py > gpurun_out/r02_run20_bench_cfg2.json 2> gpurun_out/r02_run20_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_run20_bench_cfg2.json").read().splitlines() if l.startswith("{")][0])
print(round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["clocks"], "frac", d["roofline"]["frac"], "cpu", d["cpu_baseline"])
PY
