import json
j = json.loads(open('gpurun_out/r02_s21/sweep_2m.json').read().strip().splitlines()[-1])
print('ms', round(j['ms_per_step'], 1), 'clk', j['clocks']['sm_mhz'], j['clocks']['power_w_median'])
print(json.dumps(j['check'].get('int8_vs_fp64_full_n'))[:500])
for k in j['roofline']['kernels']:
    print('    ', k['slot'], k['launches'], round(k['ms_total'] / j['steps'], 1), round(k['share_of_step'], 4), k.get('issued_int8_tops'))
