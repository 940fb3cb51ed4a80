"""Library-side fp64 roofline probe: cuBLAS DGEMM/ZGEMM + HBM copy."""
import json, torch, time
dev = torch.device("cuda:0")
def bench(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
out = []
def gemm(m, n, k, dtype, label):
    a = torch.randn(m, k, device=dev, dtype=torch.float64).to(dtype); b = torch.randn(k, n, device=dev, dtype=torch.float64).to(dtype)
    c = torch.empty(m, n, device=dev, dtype=dtype)
    ms = bench(lambda: torch.matmul(a, b, out=c))
    f = (8 if dtype.is_complex else 2) * m * n * k
    out.append({"probe": label, "m": m, "n": n, "k": k, "dtype": str(dtype), "ms": ms, "tflops": f / ms * 1e-9})
    print(json.dumps(out[-1]), flush=True)
gemm(8192, 8192, 8192, torch.float64, "dgemm_8192")
gemm(4096, 4096, 4096, torch.complex128, "zgemm_4096")
gemm(128, 4096, 128, torch.complex128, "zgemm_rhs_128x128x4096")
gemm(1152, 4096, 128, torch.complex128, "zgemm_stacked_1152x128x4096")
gemm(256, 4096, 256, torch.float64, "dgemm_realified_256x256x4096")
gemm(729, 4096, 729, torch.complex128, "zgemm_729_apply")
gemm(729, 729, 729, torch.complex128, "zgemm_729_cube")
gemm(128, 65536, 128, torch.complex128, "zgemm_rhs_128x128x65536")
a = torch.empty(1 << 28, device=dev, dtype=torch.float64); b = torch.empty_like(a)
ms = bench(lambda: b.copy_(a))
out.append({"probe": "hbm_copy_2GiB", "ms": ms, "gbs": 2 * a.numel() * 8 / ms * 1e-6}); print(json.dumps(out[-1]))
p = torch.cuda.get_device_properties(0)
out.append({"name": p.name, "sms": p.multi_processor_count, "mem_gb": p.total_memory / 2**30}); print(json.dumps(out[-1]))
json.dump(out, open("gpurun_out/torch_probe.json", "w"), indent=1)
