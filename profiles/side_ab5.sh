# config 5: decoder weight gradients beside the per-step BPTT, SM budget of the side kernels
for v in 0 100 116 84; do
if [ $v = 0 ]; then export MRSSM_SIDE_WGRAD=0; else export MRSSM_SIDE_WGRAD=1 MRSSM_STEP_SIDE_SMS=$v; fi
python bench.py --config 5 --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('side_sms=$v ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), 'loss', d['e2e']['last_loss'])"
done
