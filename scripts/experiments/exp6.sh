#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out; export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_tc_conv.py -q -m gpu --timeout 120 -k "f8c or identity" > gpurun_out/tc_f8c.log 2>&1; echo "tc f8c rc=$?"; tail -30 gpurun_out/tc_f8c.log | cut -c1-200
ERR_IMAGES=4 timeout 900 python scripts/debug_errors.py bf16x3 fp16x3 fp16_f8c > gpurun_out/errors_f8c.log 2>&1; echo "errors rc=$?"
python - <<'PY'
import json,re
for line in open("gpurun_out/errors_f8c.log"):
    m=re.match(r"(\w+) (\{.*\})", line)
    if m and m.group(1)!="oracle_fp32":
        d=json.loads(m.group(2)); print(m.group(1), "seg max %.2e rms %.2e | pose max %s rms %s flips %d" % (d["seg"]["max"], d["seg"]["rms"], ["%.2e"%p["max"] for p in d["pose"]], ["%.2e"%p["rms"] for p in d["pose"]], d["mask_flips_vs_fp64"]))
    elif not m: print(line.rstrip()[-300:])
PY
for p in bf16x3 fp16_f8c; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision $p > gpurun_out/exp_$p.json 2> gpurun_out/exp_$p.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/exp_$p.json")); print("$p", "img/s %.0f ms %.3f" % (d["value"], d["ms_per_step"]), d["roofline"]["by_class_ms_per_step"])
except Exception as e: print("$p failed", e); print(open("gpurun_out/exp_$p.err").read()[-800:])
PY
done
