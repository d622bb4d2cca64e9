#!/usr/bin/env bash
# GPU session 9 (round 2): row maxima recorded by pass 1 (single-sweep pass-2 builder); parity suite, 1M-row sample, ncu of the builders.
set -u
O=gpurun_out/r02_s9
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
grep -E "^(FAILED|ERROR)" $O/pytest.log | head
timeout 600 python bench.py --rows 1e6 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --oracle-rows 0 --no-peaks > $O/sweep_1m.json 2> $O/sweep_1m.err
echo "sweep rc=$?"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_s9/sweep_1m.json').read().strip().splitlines()[-1])
c=j['check']['int8_vs_fp64_full_n']
print('ms',round(j['ms_per_step'],1),'lml',c['lml_rel_diff'],'grad',c['grad_max_abs_diff_over_max_abs'], c.get('grad_theta_max_abs_diff_over_max_abs_theta'), 'clk', j['clocks']['sm_mhz'], j['clocks']['power_w_median'])
print(json.dumps(j['check'].get('arithmetic_audit')))
for k in j['roofline']['kernels']: print('    ',k['slot'],k['launches'],round(k['ms_total'],1),round(k['share_of_step'],4),k.get('issued_int8_tops'))
PY
CMD="python bench.py --rows 3e5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-check --no-peaks"
K='regex:k_build_phi|k_contract|k_tables|k_ozaki|k_slot_hi'
# per evaluation: k_tables, 3 x (k_slot_hi, k_build_phi_t, k_ozaki), 8 x (k_build_phi, k_ozaki, k_contract_back, k_tables[KF], k_contract_tail) = 50
ncu --set full --clock-control none --import-source on -k "$K" -s 51 -c 3 -o $O/pass1 $CMD > $O/ncu1.log 2>&1; echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k "$K" -s 60 -c 5 -o $O/pass2 $CMD > $O/ncu2.log 2>&1; echo "ncu2 rc=$?"
for f in pass1 pass2; do ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>/dev/null; done
ncu -i $O/pass2.ncu-rep --page source --csv --kernel-name regex:k_contract_back > $O/pass2_source_back.csv 2>/dev/null
ncu -i $O/pass2.ncu-rep --page source --csv --kernel-name regex:k_contract_tail > $O/pass2_source_tail.csv 2>/dev/null
ncu -i $O/pass2.ncu-rep --page source --csv --kernel-name regex:k_build_phi > $O/pass2_source_build.csv 2>/dev/null
ncu -i $O/pass1.ncu-rep --page source --csv --kernel-name regex:k_build_phi_t > $O/pass1_source_build_t.csv 2>/dev/null
du -sm $O
sz=$(du -sm $O | cut -f1); if [ "$sz" -gt 55 ]; then rm -f $O/pass1.ncu-rep; fi
sz=$(du -sm $O | cut -f1); if [ "$sz" -gt 55 ]; then rm -f $O/pass2.ncu-rep; fi
