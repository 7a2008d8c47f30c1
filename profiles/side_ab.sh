# A/B of the decoder weight gradients beside the rollout BPTT (side stream), one B200, B=1024 x T=50, bf16
for v in 1 0 1 0; do
MRSSM_SIDE_WGRAD=$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('side_wgrad=$v ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), 'loss', d['e2e']['last_loss'])"
done
