#!/bin/bash
# A/B of the time-batched MAC variants on the C3 bench workload (one B200): prints kernel, launch ms, step ms per setting.
# settings: "<threads per CTA> <tiles per CTA, 0 = by call size> <groups per synchronisation point>" or "legacy"
for cfg in "tbw" "256 0 1" "legacy"; do
  set -- $cfg
  extra=""; export BBX_TBW=0
  if [ "$1" = "legacy" ]; then extra="--tile 116"; elif [ "$1" = "tbw" ]; then export BBX_TBW=1; else export BBX_TBS_THREADS=$1 BBX_TBS_NTILE=$2 BBX_TBS_GPS=$3; fi
  python bench.py --steps ${STEPS:-50} --warmup 5 --no-cpu --no-configs --no-mimo --no-latency --no-streaming $extra 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$cfg', d['roofline']['kernel'], 'launch_ms %.4f step_ms %.4f snr %.1f frac %.3f e2e_ms %.4f' % (d['roofline']['launch_ms'], d['ms_per_step'], d['parity']['snr_db'], d['roofline']['frac'], d['e2e']['ms_per_step']))
"
done
