import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import parity
names = sys.argv[1:] or ["pyramid"]
for name in names:
    r=parity.run_case(parity.CASES[name], n_rays=30000, seed=42)
    print(name, {k:v for k,v in r.items() if k!='layers'})
    for l in r['layers']: print('   ', l)
