"""One-screen summary of a bench.py JSON line."""
import json, sys
p = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", p['value'], "ms", round(p['ms_per_step'], 4), "e2e ms", round(p['e2e']['ms_per_step'], 4), "e2e static ms", round((p['e2e'].get('static_geometry') or {}).get('ms_per_step', 0), 4), "err", p.get('error'))
if isinstance(p.get('parity'), dict):
    print("parity", {k: (v['differing_px'], v['differing_depth']) for k, v in p['parity'].items()})
r = p['roofline']; print("raster ms", round(r['kernel_ms'], 4), "front ms", round(r['frontend_kernels_ms'], 4), "frac", round(r['frac'], 4))
for k, v in p['extra'].items():
    if 'error' in v: print(k, v['error']); continue
    if 'roofline' not in v: print(k, {a: b for a, b in v.items() if isinstance(b, (int, float))}); continue
    print(k, "ms", round(v['ms_per_step'], 4), "e2e", round(v['e2e_ms_per_step'], 4), "Gpix/s", round(v['gpix_per_s'], 2), "Mtri/s", round(v['mtri_per_s'], 1),
          "frac", round(v['roofline']['frac'], 4), "raster", round(v['roofline']['raster_ms'], 4), "front", round(v['roofline']['frontend_ms'], 4),
          *(["e2e static", round(v['e2e_static_geometry']['ms_per_step'], 4)] if 'e2e_static_geometry' in v else []))
