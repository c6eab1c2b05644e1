cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
Q="--steps 40 --warmup 5 --no-cpu-baseline --no-e2e"
python bench.py $Q > gpurun_out/o_new.json 2> gpurun_out/o.err
cp prototype2_b200/libprtc.so /tmp/new.so; cp gpurun_in_old_libprtc.so prototype2_b200/libprtc.so
python bench.py $Q > gpurun_out/o_old.json 2>> gpurun_out/o.err
cp /tmp/new.so prototype2_b200/libprtc.so
python bench.py $Q > gpurun_out/o_new2.json 2>> gpurun_out/o.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/o_*.json")):
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(j["ms_per_step"],4), round(j["ms_per_step_median"],4), round(j["ms_per_step_min"],4))
PY
