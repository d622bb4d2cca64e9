#!/usr/bin/env bash
# GPU session 12 (round 2): pass-2 builder v3 (lane = column quad, coalesced packed stores); C5-shape slab budgets on one GPU.
set -u
O=gpurun_out/r02_s12
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
grep -E "^(FAILED|ERROR)" $O/pytest.log | head
timeout 600 python bench.py --rows 1e6 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --oracle-rows 0 --no-peaks > $O/sweep_1m.json 2> $O/sweep_1m.err
echo "sweep rc=$?"
for mb in 4096 2048 1024; do
  timeout 600 python bench.py --config C5 --rows 1.5e6 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --oracle-rows 0 --no-peaks --no-check --predict-rows 0 --slab-mb $mb > $O/c5_slab_$mb.json 2> $O/c5_slab_$mb.err
  echo "c5 slab $mb rc=$?"
done
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_s12/sweep_1m.json').read().strip().splitlines()[-1])
c=j['check']['int8_vs_fp64_full_n']
print('ms',round(j['ms_per_step'],1),'lml',c['lml_rel_diff'],'grad',c['grad_max_abs_diff_over_max_abs'], c.get('grad_theta_max_abs_diff_over_max_abs_theta'), 'clk', j['clocks']['sm_mhz'], j['clocks']['power_w_median'])
for k in j['roofline']['kernels']: print('    ',k['slot'],k['launches'],round(k['ms_total'],1),round(k['share_of_step'],4),k.get('issued_int8_tops'))
for mb in (4096,2048,1024):
    try:
        j=json.loads(open('gpurun_out/r02_s12/c5_slab_%d.json'%mb).read().strip().splitlines()[-1])
        k={r['slot']:r for r in j['roofline']['kernels']}
        print('C5',mb,'ms',round(j['ms_per_step'],1),'gram',k['k_gram']['launches'],round(k['k_gram']['ms_total'],1),round(k['k_gram']['issued_int8_tops']),'build_t',round(k['k_build_phi_t']['ms_total'],1),'solve',round(k['solve']['ms_total'],1),'clk',j['clocks']['sm_mhz'],j['clocks']['power_w_median'])
    except Exception as e:
        print('C5',mb,'failed',e)
PY
CMD="python bench.py --rows 3e5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-check --no-peaks"
ncu --set full --clock-control none --import-source on -k 'regex:k_build_phi' -s 16 -c 4 -o $O/builders $CMD > $O/ncu1.log 2>&1; echo "ncu1 rc=$?"
ncu -i $O/builders.ncu-rep --page raw --csv > $O/builders_raw.csv 2>/dev/null
du -sm $O
