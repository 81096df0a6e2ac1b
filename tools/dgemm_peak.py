# Calibrates the FP64 "peak" used as the roofline denominator for the score kernel: cuBLAS DGEMM via torch.
import torch, json
torch.backends.cuda.matmul.allow_tf32 = False
res = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2): (a @ b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[n] = 2 * n ** 3 / best / 1e9
    print(f"cuBLAS DGEMM n={n}: {best:.3f} ms  {res[n]:.2f} TFLOP/s")
print(json.dumps({"dgemm_tflops": res}))
