import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'ERR', e); continue
    print('==', f, 'value %.2f Gpt/s  ms/step %.1f' % (d['value'], d['ms_per_step']))
    if 'time_to_vc_tol_ms' in d:
        print('  ', d.get('details', {}).get('v_cycles'), {k: round(v, 2) for k, v in d['time_to_vc_tol_ms'].items()})
        print('   e2e %.2f Gpt/s  %.1f ms' % (d['e2e']['value'], d['e2e']['ms_per_step']), {k: round(v, 2) for k, v in d['e2e']['stages_ms'].items()})
        print('   launches', d['gpu_launches'], 'clocks', d['clocks'])
        for k, v in d['roofline']['kernels'].items():
            print('    %-36s n=%6d avg %.4f ms  %.0f GB/s' % (k, v['launches'], v['avg_ms'], v['achieved_gbs']))
        print('   roofline frac %.3f share %.3f' % (d['roofline']['frac'], d['roofline']['share_of_step']), 'cpu', d['cpu_baseline'], json.dumps(d['check'])[:1500])
    else:
        print(d['cpu_baseline'], d['ms_per_step'])
