timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2k_tests.log; cat gpurun_out/r2k_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; tail -c 400 gpurun_out/r2k_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2k_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"], d["clocks"])
print(json.dumps(d.get("config3"))[:900])
print(json.dumps(d.get("config5"))[:400])
print(d["kernels"]["attention"], d["roofline"])
PY
