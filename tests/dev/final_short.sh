#!/bin/bash
# last sanity pass of the round: full GPU suite, smoke, the bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q > $O/final2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/final2_pytest.log )
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final2_smoke.log 2>&1; echo "smoke rc=$?" >> $O/final2_smoke.log )
( timeout 600 python bench.py --steps 20 --warmup 3 > $O/final2_bench_1gpu.json 2> $O/final2_bench_1gpu.err; echo "bench rc=$?" >> $O/final2_bench_1gpu.err )
tail -3 $O/final2_pytest.log; tail -2 $O/final2_smoke.log | cut -c1-300; tail -1 $O/final2_bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final2_bench_1gpu.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'frac',round(d['roofline']['frac'],3),'clocks',d['clocks']['sm_mhz'],'e2e',round(d['e2e']['value']))
print('mdft',round(d['mdft_c3']['us_per_apply'],1),round(d['mdft_c3']['roofline']['frac'],3),'c4 us/wvl',round(d['c4_polychromatic']['us_per_wavelength_per_gpu'],1),'c5 us/plane',round(d['c5_free_space']['us_per_plane'],1),'fused_psf',round(d['fused_psf']['us_per_psf'],1))
PY
( timeout 200 python tools/bench_paths.py > gpurun_out/final2_paths.log 2>&1 ); cat gpurun_out/final2_paths.log
