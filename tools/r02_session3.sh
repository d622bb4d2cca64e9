#!/usr/bin/env bash
# GPU session 3 (round 2): reverse-mode contraction kernel + P^-1 GEMM; digit sweep of pass 2.
set -u
O=gpurun_out/r02_s3
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
grep -E "^(FAILED|ERROR)" $O/pytest.log | head
for dg in 6,7 6,6 6,5 6,4 6,3 5,4; do
  timeout 600 python bench.py --rows 1e6 --steps 2 --warmup 1 --digits $dg --no-e2e --no-cpu-baseline --oracle-rows 0 --no-peaks \
    > $O/sweep_$dg.json 2> $O/sweep_$dg.err
  echo "sweep $dg rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_s3/sweep*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'unparsable',e); continue
    c=j['check']['int8_vs_fp64_full_n']
    print(f.split('/')[-1], 'ms',round(j['ms_per_step'],1),'lml',c['lml_rel_diff'],'grad',c['grad_max_abs_diff_over_max_abs'], 'clk', j['clocks']['sm_mhz'], j['clocks']['power_w_median'])
    for k in j['roofline']['kernels']: print('    ',k['slot'],k['launches'],round(k['ms_total'],1),round(k['share_of_step'],4),k.get('issued_int8_tops'))
PY
