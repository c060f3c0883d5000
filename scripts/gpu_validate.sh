#!/bin/bash
# GPU box: the full GPU test-suite, then the four bench workloads and the reference arm; lines land in gpurun_out/final_*.log
mkdir -p gpurun_out; rm -f gpurun_out/parity_r2.log
python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/final_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench_simmim.log 2>&1; echo "simmim rc=$?"
python bench.py --workload dino --steps 10 --warmup 3 > gpurun_out/final_bench_dino.log 2>&1; echo "dino rc=$?"
python bench.py --arch vit_b --steps 10 --warmup 3 > gpurun_out/final_bench_vitb.log 2>&1; echo "vitb rc=$?"
python bench.py --workload dino --arch vit_b --batch 64 --steps 6 --warmup 3 > gpurun_out/final_bench_dino_vitb.log 2>&1; echo "dino_vitb rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference_arm.log 2>&1; echo "ref rc=$?"
for f in simmim dino vitb dino_vitb; do python - <<P
import json
d=json.loads(open("gpurun_out/final_bench_$f.log").readline())
print("$f", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "mfu", d["mfu_vs_sustained_bf16"], "gemm", d["roofline"]["frac"], "attn", d["roofline_attn"]["frac"], "ln", d["roofline_hbm"]["frac"], (d.get("torch_gpu") or {}).get("value"))
P
done
tail -c 400 gpurun_out/final_bench_reference_arm.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
