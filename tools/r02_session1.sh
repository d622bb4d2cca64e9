#!/usr/bin/env bash
# GPU session 1 (round 2): parity suite, digit-count sweep on a 1M-row C3 sample, one full C3 line.
set -u
O=gpurun_out/r02_s1
mkdir -p $O
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
nproc > $O/host.txt; free -g >> $O/host.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
for dg in 7,7 6,6 6,5 5,5 7,4; do
  timeout 600 python bench.py --rows 1e6 --steps 2 --warmup 1 --digits $dg --no-e2e --no-cpu-baseline --oracle-rows 0 \
    > $O/sweep_$dg.json 2> $O/sweep_$dg.err
  echo "sweep $dg rc=$?"
done
timeout 900 python bench.py --steps 2 --warmup 2 --power-trace $O/power_c3.json > $O/bench_c3.json 2> $O/bench_c3.err
echo "bench rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_s1/*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'unparsable',e); continue
    if 'value' not in j: continue
    r=j.get('roofline') or {}
    print(f, 'value',round(j['value'],4),'ms',round(j['ms_per_step'],1),'frac',r.get('frac'),'nongemm',r.get('non_gemm_row_kernels_share_of_step'))
    print('   check', json.dumps(j.get('check'))[:600])
    for k in r.get('kernels',[]): print('   ',k['slot'],k['launches'],round(k['ms_total'],1),round(k['share_of_step'],4),k.get('issued_int8_tops'))
PY
