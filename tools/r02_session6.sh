#!/usr/bin/env bash
# GPU session 6 (round 2): parity suite (host-evaluated kernels, device Khatri-Rao mat-vec, pipelined contraction kernel) + 1M-row C3 sample.
set -u
O=gpurun_out/r02_s6
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
grep -E "^(FAILED|ERROR)" $O/pytest.log | head
timeout 600 python bench.py --rows 1e6 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --oracle-rows 0 --no-peaks > $O/sweep_1m.json 2> $O/sweep_1m.err
echo "sweep rc=$?"
timeout 300 python tools/perf_eval.py > $O/perf_eval.txt 2>&1
tail -4 $O/perf_eval.txt
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_s6/sweep_1m.json').read().strip().splitlines()[-1])
c=j['check']['int8_vs_fp64_full_n']
print('ms',round(j['ms_per_step'],1),'lml',c['lml_rel_diff'],'grad',c['grad_max_abs_diff_over_max_abs'], 'clk', j['clocks']['sm_mhz'], j['clocks']['power_w_median'])
for k in j['roofline']['kernels']: print('    ',k['slot'],k['launches'],round(k['ms_total'],1),round(k['share_of_step'],4),k.get('issued_int8_tops'))
PY
