"""Print the interesting parts of a bench.py JSON line:  python tools/bench_summary.py gpurun_out/x.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "ms_per_step_cuda_events", "gpu_launches", "n_gpus"):
    print(k, d.get(k))
print("e2e", d.get("e2e"))
r = d.get("roofline") or {}
print("roofline", {k: v for k, v in r.items() if k in ("achieved", "frac", "kernel_ms", "traffic", "peak")})
print("phase", d.get("phase_ms"))
print("batch", d.get("batch"))
print("strong", d.get("strong"))
cb = d.get("cpu_baseline") or {}
print("cpu", {k: (v if k != "reference_sample" else {kk: vv for kk, vv in (v or {}).items() if kk not in ("what",)}) for k, v in cb.items()})
print("wall", d.get("icl_wall"))
print("wall strong", d.get("icl_wall_strong_workload"))
for k in ("host_ranked", "device_rank_option", "portable_numpy", "avx2_numpy"):
    v = d.get(k) or {}
    print(k, {kk: v.get(kk) for kk in ("value", "ms_per_step", "e2e_ms_per_step", "identical_to_oracle_with_this_numpy", "error") if kk in v})
print("clocks", d.get("clocks"))
