#!/usr/bin/env bash
# GPU session 17 (round 2): grouped rasterisation for CTA pairs, pairs only on >= 16 x 16 tiles: parity suite, C3 1M sample, C2, ncu of the pair Z GEMM.
set -u
O=gpurun_out/r02_s17
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
grep -E "^(FAILED|ERROR)" $O/pytest.log | head
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 600 python bench.py --rows 1e6 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline --no-peaks --oracle-rows 0 > $O/sweep_1m.json 2> $O/sweep_1m.err
echo "sweep rc=$?"
timeout 600 python bench.py --config C2 --steps 3 --warmup 3 --no-cpu-baseline --oracle-rows 0 > $O/c2.json 2> $O/c2.err
echo "c2 rc=$?"
python - <<'PY'
import json
for f in ('sweep_1m','c2'):
    j=json.loads(open('gpurun_out/r02_s17/%s.json'%f).read().strip().splitlines()[-1])
    print(f,'value',round(j['value'],4),'ms',round(j['ms_per_step'],1),'clk', j['clocks']['sm_mhz'], j['clocks']['power_w_median'])
    print('  ',json.dumps(j['check'].get('int8_vs_fp64_full_n'))[:500])
    for k in j['roofline']['kernels']: print('    ',k['slot'],k['launches'],round(k['ms_total'],1),round(k['share_of_step'],4),k.get('issued_int8_tops'))
PY
CMD="python bench.py --rows 3e5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-check --no-peaks"
ncu --set full --clock-control none --import-source on -k 'regex:k_ozaki' -s 30 -c 1 -o $O/k_ozaki_z $CMD > $O/ncu_b.log 2>&1; echo "ncu z rc=$?"
ncu -i $O/k_ozaki_z.ncu-rep --page raw --csv > $O/k_ozaki_z_raw.csv 2>/dev/null
du -sm $O
