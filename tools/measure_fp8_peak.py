"""What does this box sustain on an 8-bit tensor-core GEMM?  (MEASURED_PEAKS.json only has bf16.)
torch._scaled_mm (cuBLASLt) on e4m3 x e4m3 -> bf16, 8192^3, best of 10 (burst) and a 4 s loop (sustained),
with clocks and power sampled during the loop.  Library call: a yardstick, not part of the product."""
import subprocess, threading, time, json
import torch
dev = torch.device("cuda:0")
n = 8192
def bench(fn, flop):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    rows = []
    stop = threading.Event()
    def sample():
        p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
        while not stop.is_set():
            l = p.stdout.readline()
            if not l: break
            rows.append([float(x) for x in l.split(",")])
        p.terminate()
    th = threading.Thread(target=sample, daemon=True); th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); k = 0
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(20): fn()
        k += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize(); stop.set()
    sus = e0.elapsed_time(e1) / k
    busy = sorted(r[0] for r in rows)[len(rows) // 2:]
    pw = sorted(r[1] for r in rows)[len(rows) // 2:]
    return {"burst_tflops": flop / best / 1e9, "sustained_tflops": flop / sus / 1e9, "sm_mhz_median_under_load": busy[len(busy) // 2] if busy else None,
            "power_w_median_under_load": pw[len(pw) // 2] if pw else None}
a16 = torch.randn(n, n, device=dev, dtype=torch.bfloat16); b16 = torch.randn(n, n, device=dev, dtype=torch.bfloat16)
out = {"bf16_matmul_8192": bench(lambda: torch.matmul(a16, b16), 2.0 * n ** 3)}
a8 = torch.randn(n, n, device=dev).to(torch.float8_e4m3fn); b8 = torch.randn(n, n, device=dev).to(torch.float8_e4m3fn).t()
one = torch.ones((), device=dev)
try:
    out["fp8_scaled_mm_8192"] = bench(lambda: torch._scaled_mm(a8, b8, scale_a=one, scale_b=one, out_dtype=torch.bfloat16), 2.0 * n ** 3)
except Exception as e:
    out["fp8_scaled_mm_8192"] = {"error": str(e)[:200]}
print(json.dumps(out))
