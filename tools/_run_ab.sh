# attention v2 kernel first (bounded), then the NVDEC lock A/B
CB_ATTN_KERNEL=tc2 timeout 150 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attention and not agree" 2>&1 | tail -6
CB_ATTN_KERNEL=tc2 timeout 60 python tools/prof_attn.py 264
timeout 60 python tools/prof_attn.py 264
timeout 300 python -m pytest tests/test_gpu_dedup.py tests/test_gpu_decode.py tests/test_gpu_stages.py -m gpu -q 2>&1 | tail -6
for n in 12 20; do timeout 150 python bench.py --decoders $n --no-cpu-baseline --no-shots --steps 3 > gpurun_out/bench_dec$n.json 2>gpurun_out/bench_dec$n.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_dec$n.json"))
print($n, "e2e", round(d["e2e"]["value"],2), "seek", round(d["e2e_keyframe_seek"]["value"],2), d["decode_roofline"], d.get("e2e_error"))
PY
done
CB_NVDEC_CTX_LOCK=1 timeout 150 python bench.py --no-cpu-baseline --no-shots --steps 3 > gpurun_out/bench_lock.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/bench_lock.json")); print("lock e2e", d["e2e"]["value"], d["decode_roofline"])
PY
