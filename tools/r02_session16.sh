#!/usr/bin/env bash
# GPU session 16 (round 2): what would 5 digits for the Gram cost in accuracy?  INT8-vs-FP64 at full n on C2, C4 (2M rows), C5 shape (1.5M rows), C3 (2M rows).
set -u
O=gpurun_out/r02_s16
mkdir -p $O
run() { name=$1; shift; timeout 600 python bench.py "$@" --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-peaks --oracle-rows 0 --predict-rows 0 > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; }
for dg in 5,4 6,4; do
  run c2_$dg --config C2 --digits $dg
  run c4_$dg --config C4 --rows 2e6 --digits $dg
  run c5_$dg --config C5 --rows 1.5e6 --digits $dg
  run c3_$dg --config C3 --rows 2e6 --digits $dg
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_s16/*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'unparsable'); continue
    c=j['check']['int8_vs_fp64_full_n']; a=j['check'].get('arithmetic_audit')
    print(f.split('/')[-1],'ms',round(j['ms_per_step'],1),'lml',c['lml_rel_diff'],'grad',c['grad_max_abs_diff_over_max_abs'],c.get('grad_theta_max_abs_diff_over_max_abs_theta'),'audit',a and a['gram'])
PY
