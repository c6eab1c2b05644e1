cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/q_bench.json 2> gpurun_out/q.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:k_step_stream -s 3 -c 1 -f -o gpurun_out/q_kss python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>> gpurun_out/q.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/q_bench.json").read().strip().splitlines()[-1])
print(j["value"], j["ms_per_step"], j["roofline"]["frac"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["cpu_baseline"]["value"], j["gpu_launches"], j["clocks"])
PY
