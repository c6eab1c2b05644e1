cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/u_bench.json 2> gpurun_out/u.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/u_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>> gpurun_out/u.err
PRTC_CROSS_FRAME=0 ncu --set full --clock-control none --import-source on -k regex:k_xc_refresh -s 3 -c 1 -f -o gpurun_out/u_xc_top1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>> gpurun_out/u.err
PRTC_CROSS_FRAME=0 PRTC_TOP_STAGING=0 ncu --set full --clock-control none --import-source on -k regex:k_xc_refresh -s 3 -c 1 -f -o gpurun_out/u_xc_top0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>> gpurun_out/u.err
ncu --set full --clock-control none --import-source on -k regex:k_step_stream -s 3 -c 1 -f -o gpurun_out/u_kss python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>> gpurun_out/u.err
PRTC_CROSS_FRAME=0 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/u_xc0_top1.json 2>> gpurun_out/u.err
PRTC_CROSS_FRAME=0 PRTC_TOP_STAGING=0 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/u_xc0_top0.json 2>> gpurun_out/u.err
ls -la gpurun_out/u_*; tail -3 gpurun_out/u.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/u_*.json")):
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(j["ms_per_step"],4), round(j["roofline"]["frac"],4), (j.get("e2e") or {}).get("ms_per_step"))
PY
