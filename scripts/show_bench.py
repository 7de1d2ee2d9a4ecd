"""Pretty-print a bench.py JSON line: python scripts/show_bench.py gpurun_out/bench.json"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.lstrip().startswith('{')][-1])
print(f"value {d['value']/1e6:.2f} M samples/s   {d['ms_per_step']:.4f} ms/step   launches {d['gpu_launches']}   clocks {d['clocks']}")
print(f"host issue {d.get('host_issue_ms_per_step')} ms/step   kernel sum {d.get('kernel_ms_sum')} ms")
for name in ("roofline", "roofline_pool_fwd"):
    r = d.get(name)
    if r: print(f"{name:18s} {r['achieved']:.0f} GB/s  frac {r['frac']:.3f} (of 8 TB/s {r['frac_of_nominal_8TBs']:.3f})  {r['ms']*1e3:.1f} us")
print("pool only", d["pool_kernels_only"])
if d.get("e2e"): print(f"e2e {d['e2e']['value']/1e6:.2f} M samples/s  {d['e2e']['ms_per_step']:.3f} ms/step")
if d.get("cpu_baseline"): print(f"cpu {d['cpu_baseline']['value']:.0f} samples/s on {d['cpu_baseline']['cores']} cores")
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms"]):
    g = d["gemm_tensor_pipe"].get(k)
    print(f"  {k:18s} {v['ms']*1e3:8.1f} us x{v['calls_per_step']:.0f}", f"  {g['tflops']:.0f} TFLOP/s ({g['frac_of_peak']:.2f})" if g else "")
