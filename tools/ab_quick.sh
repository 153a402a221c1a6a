#!/bin/bash
# quick A/B of libbbx builds on the C3 bench workload (no ncu pass): step, MAC launch, SNR per build; "base" = in-tree build
Q="--no-streaming --no-cpu --no-latency --no-mimo --no-configs"
for v in "$@"; do
  if [ "$v" = base ]; then unset BBX_LIB; else export BBX_LIB=$PWD/bbcat-dsp_b200/variants/libbbx_$v.so; fi
  python bench.py --steps ${STEPS:-200} --warmup 5 $Q 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-10s step_ms %.4f mac_ms %.4f rest_us %.1f snr %.1f' % ('$v', d['ms_per_step'], d['roofline']['launch_ms'], 1e3*(d['ms_per_step']-d['roofline']['launch_ms']), d['parity']['snr_db']))"
done
