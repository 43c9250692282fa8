import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ['value', 'ms_per_step', 'gpu_launches', 'clocks']}, 'e2e', d['e2e'], 'single_image', d.get('single_image'), 'frac', round(d['roofline']['frac'], 4), 'TF', round(d['roofline']['achieved'], 1), 'gemm', d['roofline'].get('dominant_kernel'))
kb = d.get('kernel_breakdown_ms_per_forward')
if kb:
    tot = sum(v['ms'] for v in kb.values())
    print('forward of 40 samples total ms', round(tot, 2))
    for k, v in kb.items():
        print(f"{k:20s} {v['ms']:9.3f} ms {v['launches']:4d}  {100*v['ms']/tot:5.1f}%")
