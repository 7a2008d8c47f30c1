# 8-GPU lines: config 3 (headline) and config 5, one rank per GPU (torchrun), plus the 1-GPU line of the same box
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_bf16_b1024_n8.json 2> gpurun_out/n8.err
tail -c 300 gpurun_out/n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --config 5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_bf16_config5_n8.json 2> gpurun_out/n8c5.err
tail -c 300 gpurun_out/n8c5.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_bf16_b1024_n1_samebox.json 2>/dev/null
for f in gpurun_out/r02_bench_bf16_b1024_n8.json gpurun_out/r02_bench_bf16_config5_n8.json gpurun_out/r02_bench_bf16_b1024_n1_samebox.json; do tail -1 $f | cut -c1-330; done
