cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
Q="--steps 40 --warmup 5 --no-cpu-baseline --no-e2e"
cp prototype2_b200/libprtc.so /tmp/new.so
for v in new old div3 new old div3; do
  case $v in new) cp /tmp/new.so prototype2_b200/libprtc.so;; old) cp gpurun_in_old_libprtc.so prototype2_b200/libprtc.so;; div3) cp gpurun_in_div3_libprtc.so prototype2_b200/libprtc.so;; esac
  python bench.py $Q 2>> gpurun_out/p.err | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(j['ms_per_step'],4), round(j['ms_per_step_median'],4), round(j['ms_per_step_min'],4))"
done
