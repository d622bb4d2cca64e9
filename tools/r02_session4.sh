#!/usr/bin/env bash
# GPU session 4 (round 2): new digit defaults (6, 4, 6): parity suite, ncu --set full of every hot kernel, launch list, C3 / C2 / C4 lines.
set -u
O=gpurun_out/r02_s4
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
grep -E "^(FAILED|ERROR)" $O/pytest.log | head
CMD="python bench.py --rows 3e5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-check --no-peaks"
$CMD > $O/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:k_build_phi|k_contract_rows|k_tables|k_ozaki|k_slot_hi' -s 34 -c 34 \
    -o $O/hot_kernels $CMD > $O/ncu.log 2>&1
echo "ncu rc=$?"
CMD2="python bench.py --rows 4e5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-check --no-peaks"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv $CMD2 > $O/ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 900 python bench.py --steps 2 --warmup 3 --power-trace $O/power_c3.json > $O/bench_c3.json 2> $O/bench_c3.err
echo "bench C3 rc=$?"
timeout 600 python bench.py --config C2 --steps 3 --warmup 3 > $O/bench_c2.json 2> $O/bench_c2.err
echo "bench C2 rc=$?"
timeout 900 python bench.py --config C4 --steps 2 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err
echo "bench C4 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_s4/bench*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'unparsable',e); continue
    r=j.get('roofline') or {}
    print(f, 'value',round(j['value'],4),'e2e',j['e2e'] and round(j['e2e']['value'],4),'ms',round(j['ms_per_step'],1),'frac',r.get('frac'),'nongemm',r.get('non_gemm_row_kernels_share_of_step'), j['clocks'].get('sm_mhz'), j['clocks'].get('power_w_median'))
    print('   check', json.dumps(j.get('check'))[:900])
    print('   cpu', json.dumps(j.get('cpu_baseline'))[:400])
    for k in r.get('kernels',[]): print('   ',k['slot'],k['launches'],round(k['ms_total'],1),round(k['share_of_step'],4),k.get('issued_int8_tops'))
PY
