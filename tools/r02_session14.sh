#!/usr/bin/env bash
# GPU session 14 (round 2): CTA-pair GEMM as the default arithmetic: parity suite, 1M-row C3 sample and a C5-shape sample with the full checks.
set -u
O=gpurun_out/r02_s14
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
grep -E "^(FAILED|ERROR)" $O/pytest.log | head
timeout 600 python bench.py --rows 1e6 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-peaks > $O/sweep_1m.json 2> $O/sweep_1m.err
echo "sweep rc=$?"
timeout 600 python bench.py --config C5 --rows 1.5e6 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-peaks --predict-rows 20000 > $O/c5_sample.json 2> $O/c5_sample.err
echo "c5 rc=$?"
python - <<'PY'
import json
for f in ('sweep_1m','c5_sample'):
    j=json.loads(open('gpurun_out/r02_s14/%s.json'%f).read().strip().splitlines()[-1])
    print(f,'ms',round(j['ms_per_step'],1),'clk', j['clocks']['sm_mhz'], j['clocks']['power_w_median'])
    print('  ',json.dumps(j['check'])[:1200])
    print('  ',json.dumps(j.get('predict')))
    for k in j['roofline']['kernels']: print('    ',k['slot'],k['launches'],round(k['ms_total'],1),round(k['share_of_step'],4),k.get('issued_int8_tops'))
PY
