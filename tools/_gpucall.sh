for m in 0 1 2 0 1 2; do
  BBX_TB_MODE=$m python bench.py --steps 300 --warmup 5 --no-cpu --no-latency --no-mimo --no-streaming > gpurun_out/ab_$m.json 2>gpurun_out/ab_$m.err
  python - <<PY
import json
j=json.loads(open('gpurun_out/ab_$m.json').read().strip().splitlines()[-1])
print('mode $m', j['ms_per_step'], j['roofline']['launch_ms'], j['e2e']['ms_per_step'])
PY
done
