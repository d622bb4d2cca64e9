#!/usr/bin/env bash
# GPU session 19 (round 2): final code: parity suite, smoke(), default bench invocation of the driver (python bench.py), reference arm.
set -u
O=gpurun_out/r02_s19
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
grep -E "^(FAILED|ERROR)" $O/pytest.log | head
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
/usr/bin/time -v python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?"
grep -E "Elapsed|Maximum resident" $O/bench_default.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_s19/bench_default.json').read().strip().splitlines()[-1])
print('value',j['value'],'e2e',j['e2e'],'ms',j['ms_per_step'],'launches',j['gpu_launches'],'clocks',j['clocks'])
print('roofline',{k:j['roofline'][k] for k in ('bound','achieved','peak','unit','frac')}, j['roofline']['traffic'])
print('cpu',json.dumps(j['cpu_baseline'])[:300])
PY
