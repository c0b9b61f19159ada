# same-box A/B of the headline self-play bench: programmatic dependent launch on / off (2 timed plies of 4096 games x 800 sims)
mkdir -p gpurun_out
export XQ_BENCH_NO_CONFIGS3=1 XQ_BENCH_NO_TRAIN=1
for pdl in 1 0 1 0; do
  XQ_NET_PDL=$pdl timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']
print('XQ_NET_PDL=$pdl', round(d['value']), 'sims/s', round(d['ms_per_step'],1), 'ms/ply', 'frac', round(r['frac'],4), 'fwd isolated', round(r['forward_ms_isolated'],4), 'sustained', round(r['forward_ms_back_to_back_400ms'],4), 'sm_mhz', d['clocks']['sm_mhz'])"
done | tee gpurun_out/r2_pdl_ab.txt
