cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "top_staging or cross_frame or streamed or c3 or jacobi" 2>&1 | tail -4
PRTC_CROSS_FRAME=0 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_xc_refresh -c 8 --csv --log-file gpurun_out/v_ncu_xc_top1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
PRTC_CROSS_FRAME=0 PRTC_TOP_STAGING=0 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_xc_refresh -c 8 --csv --log-file gpurun_out/v_ncu_xc_top0.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
PRTC_CROSS_FRAME=0 ncu --set full --clock-control none --import-source on -k regex:k_xc_refresh -s 3 -c 1 -f -o gpurun_out/v_xc_top1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>> gpurun_out/v.err
PRTC_CROSS_FRAME=0 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/v_xc0_top1.json 2>> gpurun_out/v.err
PRTC_CROSS_FRAME=0 PRTC_TOP_STAGING=0 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/v_xc0_top0.json 2>> gpurun_out/v.err
grep -h "k_xc_refresh" gpurun_out/v_ncu_xc_top1.csv | awk -F'","' '{print "top1", $NF}' | tail -8
grep -h "k_xc_refresh" gpurun_out/v_ncu_xc_top0.csv | awk -F'","' '{print "top0", $NF}' | tail -8
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/v_*.json")):
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(j["ms_per_step"],4), round(j["roofline"]["frac"],4))
PY
