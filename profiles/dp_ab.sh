# 2-GPU A/B of the gradient exchange: overlapped buckets vs one all-reduce after backward vs overlapped with fewer NCCL CTAs
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$1', 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), 'strong', round(d['strong_scaling']['ms_per_step'], 3))"; }
MRSSM_DP_OVERLAP=1 run overlap
MRSSM_DP_OVERLAP=0 run single_allreduce
MRSSM_DP_OVERLAP=1 NCCL_MAX_CTAS=4 run overlap_maxctas4
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('1gpu ms/step', round(d['ms_per_step'], 3))"
