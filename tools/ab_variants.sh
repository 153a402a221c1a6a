#!/bin/bash
# A/B of libbbx builds on the C3 workload: per build the bench step time / MAC launch / in-run SNR and the ncu launch list
# (cold, serialised) of one 64-block call.  usage: tools/ab_variants.sh <name>...   ("base" = the in-tree libbbx.so)
Q="--no-streaming --no-cpu --no-latency --no-mimo --no-configs"
for v in "$@"; do
  if [ "$v" = base ]; then unset BBX_LIB; else export BBX_LIB=$PWD/bbcat-dsp_b200/variants/libbbx_$v.so; fi
  python bench.py --steps ${STEPS:-300} --warmup 5 $Q > gpurun_out/q_$v.json 2> gpurun_out/q_$v.err
  python - $v <<'P'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/q_%s.json'%v).read().strip().splitlines()[-1])
    print('%-8s step_ms %.4f mac_ms %.4f rest_us %.1f e2e_ms %.4f snr %.1f' % (v, d['ms_per_step'], d['roofline']['launch_ms'], 1e3*(d['ms_per_step']-d['roofline']['launch_ms']), d['e2e']['ms_per_step'], d['parity']['snr_db']))
except Exception as e:
    print(v, 'bench failed', e); print(open('gpurun_out/q_%s.err'%v).read()[-800:])
P
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_$v.csv python tools/profile_c3.py 3 > gpurun_out/ncu_$v.log 2>&1
  python - $v <<'P'
import csv,re,sys
v=sys.argv[1]
rows=list(csv.reader(open('gpurun_out/l_%s.csv'%v)))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
seq=[(re.sub(r'\(.*','',r[4].split('bbx::')[1] if 'bbx::' in r[4] else r[4][:30]),float(r[-1])/1000) for r in rows[hi+1:] if len(r)>=15]
print('   ', ' '.join('%s %.1f'%(n.replace('k_',''),t) for n,t in seq[-6:]))
P
done
