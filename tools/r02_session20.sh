#!/usr/bin/env bash
# GPU session 20 (round 2): final code (blocked launch order of the Gram's pair tiles): parity suite, smoke(), the driver's default
# bench invocation with its wall time, a C5-shape sample, ncu of the Gram launch.
set -u
O=gpurun_out/r02_s20
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
grep -E "^(FAILED|ERROR)" $O/pytest.log | head
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
T0=$(date +%s); python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$? wall=$(( $(date +%s) - T0 )) s"
timeout 600 python bench.py --config C5 --rows 1.5e6 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-peaks --oracle-rows 0 --predict-rows 0 > $O/c5_sample.json 2> $O/c5_sample.err
echo "c5 sample rc=$?"
python tools/r02_session20_summary.py
CMD="python bench.py --rows 3e5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-check --no-peaks"
ncu --set full --clock-control none -k 'regex:k_ozaki' -s 20 -c 1 -o $O/k_ozaki_gram $CMD > $O/ncu_a.log 2>&1; echo "ncu gram rc=$?"
ncu -i $O/k_ozaki_gram.ncu-rep --page raw --csv > $O/k_ozaki_gram_raw.csv 2>/dev/null
du -sm $O
