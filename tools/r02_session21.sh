#!/usr/bin/env bash
# GPU session 21 (round 2): pass-1 slabs in whole builder waves (37888 rows at p = 4096): parity suite, C3 2M-row sample, ncu of the Gram launch.
set -u
O=gpurun_out/r02_s21
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
grep -E "^(FAILED|ERROR)" $O/pytest.log | head
timeout 600 python bench.py --rows 2e6 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline --no-peaks --oracle-rows 0 > $O/sweep_2m.json 2> $O/sweep_2m.err
echo "sweep rc=$?"
python tools/r02_session21_summary.py
CMD="python bench.py --rows 3e5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-check --no-peaks"
ncu --set full --clock-control none -k 'regex:k_ozaki' -s 18 -c 1 -o $O/k_ozaki_gram $CMD > $O/ncu_a.log 2>&1; echo "ncu gram rc=$?"
ncu -i $O/k_ozaki_gram.ncu-rep --page raw --csv > $O/k_ozaki_gram_raw.csv 2>/dev/null
du -sm $O
