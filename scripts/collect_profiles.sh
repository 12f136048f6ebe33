#!/bin/bash
# copies the round-2 evidence from gpurun_out/ (scratch) into profiles/ (tracked) under stable names
cd "$(dirname "$0")/.."
cp_if() { [ -f "gpurun_out/$1" ] && cp "gpurun_out/$1" "profiles/$2"; }
last_json() { [ -f "gpurun_out/$1" ] && tail -1 "gpurun_out/$1" > "profiles/$2"; }
cp_if r2c1_probe.log r02_probe_elements_before.log
cat gpurun_out/r2c4_tune_2d_p2.log gpurun_out/r2c4_tune_3d_p1.log gpurun_out/r2c4_tune_3d_p2.log gpurun_out/r2c4_tune_3d_p3.log > profiles/r02_tune_sweep1.log 2>/dev/null
cp_if r2c5_tune.log r02_tune_sweep2.log
cp_if r2c6_tune.log r02_tune_sweep3.log
cp_if r2c7_tune.log r02_tune_sweep4_final_variants.log
cp_if r2c9_pdl.log r02_pdl_modes.log
last_json r2c3_n2_fused.json r02_bench_n2_fused_exchange.json
last_json r2c3_n2_split.json r02_bench_n2_two_stream_exchange.json
last_json r2c7_bench.json r02_bench_n1.json
last_json r2c7_box3d.json r02_bench_box3d_n1_pdl.json
last_json r2c7_box3d_nopdl.json r02_bench_box3d_n1_nopdl.json
last_json r2c4_bench_ref.json r02_bench_reference_arm.json
last_json r2c8_n8.json r02_scale_n8.json
last_json r2c12_n8.json r02_scale_n8_batched_push.json
last_json r2c12_n1.json r02_scale_n1_same_box_batched_push.json
last_json r2c11_n2.json r02_bench_n2_batched_push.json
cat gpurun_out/r2c10_small.log gpurun_out/r2c11_small.log > profiles/r02_small_problem.log 2>/dev/null
last_json r2c8_n4.json r02_scale_n4.json
last_json r2c8_n1.json r02_scale_n1_same_box.json
cp_if r2c8_topo.txt r02_scale_topology.txt
[ -f gpurun_out/r2c2_n2_launches.csv ] && python scripts/summarize_ncu.py launches gpurun_out/r2c2_n2_launches.csv profiles/r02_launches_n2_rank0.md > /dev/null
[ -f gpurun_out/r2_launches.csv ] && python scripts/summarize_ncu.py launches gpurun_out/r2_launches.csv profiles/r02_launches_bench.md > /dev/null
for n in 2d_p2 3d_p1 3d_p2 3d_p3 2d_p4 2d_p3; do cp_if r2_full_$n.md r02_full_$n.md; done
for f in gpurun_out/r2_sanitizer_*.log; do [ -f "$f" ] && cp "$f" profiles/$(basename "$f" | sed 's/^r2_/r02_/'); done

last_json r2c29_bench.json r02_bench_n1_final.json
last_json r2c29_bench_ref.json r02_bench_reference_arm_final.json
cp_if r2c29_pytest.log r02_pytest_gpu.log
ls profiles | grep r02
