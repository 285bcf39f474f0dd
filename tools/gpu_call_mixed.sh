#!/bin/bash
TAG=${1:-r02s}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${TAG}_build.log 2>&1
timeout 300 python tools/mixed_decode_probe.py 2> gpurun_out/${TAG}_probe.err | tee gpurun_out/${TAG}_probe.json
timeout 600 python -m pytest tests -m gpu -x -q -k "stages or decode" > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-library --no-shots --e2e-steps 1 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/'+"${TAG}"+'_bench.json').read().strip().splitlines()[-1])
print('e2e',d['e2e']['value'],'c5',d['secondary']['c5_mix']['e2e_clips_per_sec'],'iv2',d['secondary']['iv2_tubes']['e2e_clips_per_sec'],'c4',d['secondary']['c4_shape']['e2e_clips_per_sec'])
PY
