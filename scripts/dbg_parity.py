import sys; sys.path.insert(0,'tests')
import parity
for name in ["partial_prob","stoch_config5","two_layer_config4","two_populations"]:
    r=parity.run_case(parity.CASES[name], n_rays=30000, seed=42)
    print(name, {k:v for k,v in r.items() if k!='layers'})
    for l in r['layers']: print('   ', l)
