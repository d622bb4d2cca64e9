#!/usr/bin/env bash
# GPU session 10 (round 2): tail kernel at 12 warps/SM; ncu of the two slab builders and the tail only.
set -u
O=gpurun_out/r02_s10
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "round2 or api or native" > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 600 python bench.py --rows 1e6 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --oracle-rows 0 --no-peaks > $O/sweep_1m.json 2> $O/sweep_1m.err
echo "sweep rc=$?"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_s10/sweep_1m.json').read().strip().splitlines()[-1])
c=j['check']['int8_vs_fp64_full_n']
print('ms',round(j['ms_per_step'],1),'lml',c['lml_rel_diff'],'grad',c['grad_max_abs_diff_over_max_abs'], c.get('grad_theta_max_abs_diff_over_max_abs_theta'), 'clk', j['clocks']['sm_mhz'], j['clocks']['power_w_median'])
for k in j['roofline']['kernels']: print('    ',k['slot'],k['launches'],round(k['ms_total'],1),round(k['share_of_step'],4),k.get('issued_int8_tops'))
PY
CMD="python bench.py --rows 3e5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-check --no-peaks"
ncu --set full --clock-control none --import-source on -k 'regex:k_build_phi' -s 30 -c 3 -o $O/builders $CMD > $O/ncu1.log 2>&1; echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:k_contract_tail' -s 20 -c 1 -o $O/tail $CMD > $O/ncu2.log 2>&1; echo "ncu2 rc=$?"
for f in builders tail; do ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>/dev/null; done
ncu -i $O/builders.ncu-rep --page source --csv --kernel-name 'regex:k_build_phi<' > $O/source_build.csv 2>/dev/null
ncu -i $O/builders.ncu-rep --page source --csv --kernel-name 'regex:k_build_phi_t' > $O/source_build_t.csv 2>/dev/null
ncu -i $O/tail.ncu-rep --page source --csv > $O/source_tail.csv 2>/dev/null
du -sm $O
