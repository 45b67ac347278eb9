#!/bin/bash
# A/B runs of the fixed-point E-step forms and build variants at config 3 (kernel development; see tools/exp_build_variant.sh)
# usage: tools/lk_sweep.sh [tile costs ...] [-- variant suffixes ...]
run() { # name, env...
  name=$1; shift
  env "$@" timeout 120 python bench.py --steps 100 --warmup 20 --quick 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('$name', 'ms/step %.4f' % d['ms_per_step'], 'e_us %.1f' % r['e_step']['us'], 'bayes_us %.1f' % r['e_step']['bayes_us'], 'm_us %.1f' % r['m_step']['us'], 'e2e %.0f' % d['e2e']['value'])
"
}
L=$PWD/scsplit_b200
timeout 200 python -m pytest tests/test_gpu_qpath.py -x -q -m gpu 2>&1 | tail -2
run default A=1
run groups SCS_Q_GROUPS=1
variants=0
for v in "$@"; do
  if [ "$v" = "--" ]; then variants=1; continue; fi
  if [ $variants = 1 ]; then run $v SCSPLIT_B200_LIB=$L/libscsplit_b200_$v.so; else run cost$v SCS_LK_TILE_COST=$v; fi
done
