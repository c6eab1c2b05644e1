cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
Q="--steps 30 --warmup 5 --no-cpu-baseline --no-e2e"
python bench.py --steps 40 --warmup 5 > gpurun_out/t_default.json 2> gpurun_out/t.err
PRTC_TUNE=0x40000000 PRTC_TIGHT_REJECT=0 python bench.py $Q > gpurun_out/t_v9.json 2>> gpurun_out/t.err
PRTC_TIGHT_REJECT=0 python bench.py $Q > gpurun_out/t_trips_notight.json 2>> gpurun_out/t.err
PRTC_TUNE=0x04000000 python bench.py $Q > gpurun_out/t_trips_e4.json 2>> gpurun_out/t.err
PRTC_TUNE=0x08000000 python bench.py $Q > gpurun_out/t_trips_e8.json 2>> gpurun_out/t.err
PRTC_TUNE=0x10000000 python bench.py $Q > gpurun_out/t_trips_e16.json 2>> gpurun_out/t.err
PRTC_TUNE=0x18000000 python bench.py $Q > gpurun_out/t_trips_e24.json 2>> gpurun_out/t.err
PRTC_TUNE=0x10000000 PRTC_TIGHT_REJECT=0 python bench.py $Q > gpurun_out/t_trips_e16_notight.json 2>> gpurun_out/t.err
PRTC_TOP_STAGING=0 python bench.py $Q > gpurun_out/t_top0.json 2>> gpurun_out/t.err
PRTC_CROSS_FRAME=0 python bench.py $Q > gpurun_out/t_xc0_top1.json 2>> gpurun_out/t.err
PRTC_CROSS_FRAME=0 PRTC_TOP_STAGING=0 python bench.py $Q > gpurun_out/t_xc0_top0.json 2>> gpurun_out/t.err
PRTC_CROSS_FRAME=0 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_xc_refresh -c 8 --csv --log-file gpurun_out/t_ncu_xc_top1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
PRTC_CROSS_FRAME=0 PRTC_TOP_STAGING=0 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_xc_refresh -c 8 --csv --log-file gpurun_out/t_ncu_xc_top0.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
for f in gpurun_out/t_*.json; do python - "$f" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(j["ms_per_step"],4), round(j.get("ms_per_step_median",0),4), round(j["roofline"]["frac"],4), j.get("e2e",{}).get("ms_per_step"), j["roofline"]["per_query"].get("exact_tests_per_query"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
grep -h "k_xc_refresh" gpurun_out/t_ncu_xc_top1.csv | awk -F'","' '{print "top1", $NF}' | tail -8
grep -h "k_xc_refresh" gpurun_out/t_ncu_xc_top0.csv | awk -F'","' '{print "top0", $NF}' | tail -8
tail -5 gpurun_out/t.err
