#!/usr/bin/env bash
# Round-2 evidence run on ONE B200: parity suite, the four 1-GPU bench lines (C3 with power trace, C2, C4, reference arm), launch list,
# ncu --set full of every hot kernel (exported to CSV on the box).
set -u
O=gpurun_out/r02_final1c
mkdir -p $O
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
nproc > $O/host.txt; free -g >> $O/host.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -2 $O/pytest.log
timeout 1200 python bench.py --steps 3 --warmup 3 --power-trace $O/power_c3.json > $O/bench_c3_n1.json 2> $O/bench_c3_n1.err
echo "bench C3 rc=$?"
timeout 600 python bench.py --config C2 --steps 3 --warmup 3 > $O/bench_c2_n1.json 2> $O/bench_c2_n1.err
echo "bench C2 rc=$?"
timeout 900 python bench.py --config C4 --steps 3 --warmup 3 > $O/bench_c4_n1.json 2> $O/bench_c4_n1.err
echo "bench C4 rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
echo "reference arm rc=$?"
timeout 900 python bench.py --gemm fp64 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --oracle-rows 0 > $O/bench_c3_n1_fp64.json 2> $O/bench_c3_n1_fp64.err
echo "bench C3 fp64 rc=$?"
CMD="python bench.py --rows 4e5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-check --no-peaks"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1
echo "launch list rc=$?"
CMD="python bench.py --rows 3e5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-check --no-peaks"
# the warm-up evaluation (+ its audit) launches 20 k_ozaki (10 Gram slabs, 8 pass-2 slabs, 2 audit), 22 builders, 20 contraction kernels:
# skip them and capture the timed evaluation's first launches
ncu --set full --clock-control none --import-source on -k 'regex:k_ozaki' -s 20 -c 1 -o $O/k_ozaki_gram $CMD > $O/ncu_a.log 2>&1; echo "ncu gram rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:k_ozaki' -s 30 -c 1 -o $O/k_ozaki_z $CMD > $O/ncu_b.log 2>&1; echo "ncu z rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:k_build_phi' -s 22 -c 1 -o $O/builder_pass1 $CMD > $O/ncu_c.log 2>&1; echo "ncu builder 1 rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:k_build_phi' -s 32 -c 1 -o $O/builder_pass2 $CMD > $O/ncu_c2.log 2>&1; echo "ncu builder 2 rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:k_contract' -s 20 -c 2 -o $O/contract $CMD > $O/ncu_d.log 2>&1; echo "ncu contract rc=$?"
ncu --set full --clock-control none -k 'regex:k_tables|k_slot_hi|k_topk|k_potf2_inv' -s 14 -c 6 -o $O/small $CMD > $O/ncu_e.log 2>&1; echo "ncu small rc=$?"
for f in k_ozaki_gram k_ozaki_z builder_pass1 builder_pass2 contract small; do ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>/dev/null; done
du -sm $O
for f in small contract builder_pass1 builder_pass2 k_ozaki_z k_ozaki_gram; do
  sz=$(du -sm $O | cut -f1); if [ "$sz" -gt 58 ]; then rm -f $O/$f.ncu-rep; fi
done
du -sm $O
