#!/bin/bash
# quick loop: parity tests of the normals / Laplacian kernels, irregular mesh, headline
timeout 300 python -m pytest tests/test_gpu_apps.py tests/test_multi.py -m gpu -x -q 2>&1 | tail -2
timeout 400 python scripts/bench_irregular.py 2237 1024 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:(round(d[k]['ms'],4), round(d[k]['hbm_frac'],3), d[k].get('max_rel_err', d[k].get('max_abs_err'))) for k in ('VV','VF','VN','LAP')})
"
timeout 300 python bench_configs.py --only queries 2>/dev/null | python -c "
import json,sys
q=json.loads(sys.stdin.read().strip().splitlines()[-1])['consume_and_normals_on_lloyd_patches']
print('lloyd icosphere', {k:(round(v['ms'],4), round(v['hbm_frac'],3)) for k,v in q.items() if isinstance(v,dict)}, q['parity_ok'])
"
timeout 300 python bench.py --sub none --steps 100 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('headline', d['ms_per_step'], {k:round(v['ms'],4) for k,v in d['kernels'].items()}, d['clocks']['sm_mhz'])
"
