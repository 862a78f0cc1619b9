"""What a pure store stream reaches on this GPU: cudaMemset (torch zero_) and a fill kernel
over 10.24 GB, CUDA events, best of 10 -- the ceiling for `outer_kernel` (config 4's d_x)."""
import json, torch
n = 1_280_000_000
a = torch.empty(n, dtype=torch.float64, device="cuda")
def best(f, reps=10):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)
for name, f in (("memset", lambda: a.zero_()), ("fill", lambda: a.fill_(1.5))):
    f(); torch.cuda.synchronize()
    ms = best(f)
    print(json.dumps({"op": name, "bytes": n * 8, "ms": round(ms, 4), "GBps": round(n * 8 / ms / 1e6, 1)}))
