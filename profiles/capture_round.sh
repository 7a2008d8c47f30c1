set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r01_bench_final.json 2> gpurun_out/r01_bench_final.err
tail -c 400 gpurun_out/r01_bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r01_launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_final.log 2>&1
tail -2 gpurun_out/ncu_launch_final.log
ncu --set full --clock-control none -k regex:'rollout_tc|plane_wgrad|plane_fwd' --launch-skip 42 --launch-count 21 --export gpurun_out/r01_full_final python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_full_final.log 2>&1
tail -2 gpurun_out/ncu_full_final.log
ncu -i gpurun_out/r01_full_final.ncu-rep --page raw --csv > gpurun_out/r01_full_final_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -8
rm -f gpurun_out/r01_full_final.ncu-rep
