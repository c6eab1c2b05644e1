export PRTC_BENCH_FAKE_WORLD=8
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --chain 0"
$B > /dev/null 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step_pool -s 4 -c 1 -f -o gpurun_out/shard_plain $B > gpurun_out/ncu_p.log 2>&1
tail -1 gpurun_out/ncu_p.log
