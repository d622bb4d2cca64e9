#!/usr/bin/env bash
# GPU session 2 (round 2): parity suite with the even-digit fix, digit sweep, ncu --set full of the row kernels and k_ozaki.
set -u
O=gpurun_out/r02_s2
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for dg in 6,6 6,7 5,7 7,6 5,6 6,5 4,7; do
  timeout 600 python bench.py --rows 1e6 --steps 2 --warmup 1 --digits $dg --no-e2e --no-cpu-baseline --oracle-rows 0 --no-peaks \
    > $O/sweep_$dg.json 2> $O/sweep_$dg.err
  echo "sweep $dg rc=$?"
done
CMD="python bench.py --rows 3e5 --steps 1 --warmup 1 --digits 6,6 --no-e2e --no-cpu-baseline --no-check --no-peaks"
$CMD > $O/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:k_build_phi|k_contract|k_dtables|k_tables|k_ozaki|k_slot_hi' -s 42 -c 16 \
    -o $O/rowkernels $CMD > $O/ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_s2/sweep*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'unparsable',e); continue
    c=j['check']['int8_vs_fp64_full_n']
    print(f.split('/')[-1], 'ms',round(j['ms_per_step'],1),'lml',c['lml_rel_diff'],'grad',c['grad_max_abs_diff_over_max_abs'])
PY
