timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -15
for g in 16 8 32; do
HG_AP_G=$g timeout 200 python bench.py --steps 5 --warmup 3 --cpu-sample 0 --no-e2e > gpurun_out/dbg$g.json 2> gpurun_out/dbg$g.err
tail -2 gpurun_out/dbg$g.err
python - <<P
import json
d=json.loads(open("gpurun_out/dbg$g.json").read().strip().splitlines()[-1])
print("G=$g", d["ms_per_step"], d["roofline"]["phases_ms"], d["parity"])
P
done
