# The exact commands behind the committed profiles/r02_* files (run on one B200 through gpurun; bash profiles/capture_round.sh).
# A number printed by a run under ncu is never a bench value: the bench lines come from the plain runs.
set -x
R=r02
python bench.py --steps 5 --warmup 3 > gpurun_out/${R}_bench_bf16_b1024.json 2> gpurun_out/${R}_bench.err
tail -c 300 gpurun_out/${R}_bench.err
python bench.py --config 5 --steps 3 --warmup 3 > gpurun_out/${R}_bench_bf16_config5.json 2> gpurun_out/${R}_bench5.err
tail -c 300 gpurun_out/${R}_bench5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${R}_launches_bf16_b1024.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
# (side stream off for this capture only: ncu serialises kernels in launch order, match_traffic.py aligns them with the in-place order)
MRSSM_SIDE_WGRAD=0 ncu --set full --clock-control none -k regex:'rollout_tc|plane_wgrad|plane_fwd' --launch-skip 48 --launch-count 24 --export gpurun_out/${R}_full python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/${R}_full.ncu-rep --page raw --csv > gpurun_out/${R}_full_raw.csv 2>/dev/null
python profiles/time_shipped_yaml.py > gpurun_out/${R}_shipped_yaml.json 2> gpurun_out/shipped.err
tail -c 600 gpurun_out/${R}_shipped_yaml.json
ls -la gpurun_out/ | tail -8
rm -f gpurun_out/${R}_full.ncu-rep
