import json, sys
for path in sys.argv[1:]:
    line = [l for l in open(path).read().splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    r = d.get("roofline", {})
    print(f"{path}: n_gpus={d['n_gpus']} ms/step={d['ms_per_step']:.3f} value={d['value']:.0f} {d['unit']} | e2e={d['e2e']['value']:.0f} "
          f"({d['e2e'].get('ms_per_step', 0):.3f} ms/step) | gemm {r.get('achieved')} TF/s frac={r.get('frac')} | clocks={d.get('clocks')} "
          f"| launches={d.get('gpu_launches')} | cpu={d.get('cpu_baseline', {}).get('value')}")
