"""Scratch GPU check for the tcgen05 GEMM: numerics vs torch fp32 and event timing."""
import ctypes as C, json, sys
import torch
sys.path.insert(0, ".")
from manual_whisper_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
torch.manual_seed(0)
res = {}
def run(M, N, K, bias=False, gelu=False, resid=False, out_f32=False):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev) if bias else None
    r = torch.randn(M, N, device=dev) if resid else None
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if out_f32 else torch.bfloat16)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.mw_gemm_h16(a.data_ptr(), w.data_ptr(), b.data_ptr() if bias else None, r.data_ptr() if resid else None,
                                out.data_ptr(), M, N, K, int(gelu), int(out_f32), st), "gemm")
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    if bias: ref = ref + b
    if gelu: ref = torch.nn.functional.gelu(ref)
    if resid: ref = ref + r
    err = (out.float() - ref).abs().max().item()
    return err, ref.abs().max().item()
shapes = [(128, 128, 64), (128, 256, 64), (128, 256, 128), (256, 256, 512), (1500, 1280, 1280), (3000, 384, 240), (200, 1152, 384), (4096, 5120, 1280), (4096, 1280, 5120)]
for (M, N, K) in shapes:
    res[f"plain_{M}x{N}x{K}"] = run(M, N, K)
res["bias_gelu"] = run(1500, 5120, 1280, bias=True, gelu=True)
res["bias_resid_f32"] = run(1500, 1280, 5120, bias=True, resid=True, out_f32=True)
res["bias_f32"] = run(777, 384, 384, bias=True, out_f32=True)
# timing
def bench(M, N, K, gelu=False, iters=20):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev); out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3): lib.mw_gemm_h16(a.data_ptr(), w.data_ptr(), b.data_ptr(), None, out.data_ptr(), M, N, K, int(gelu), 0, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): lib.mw_gemm_h16(a.data_ptr(), w.data_ptr(), b.data_ptr(), None, out.data_ptr(), M, N, K, int(gelu), 0, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    e0.record()
    for _ in range(iters): torch.matmul(a, w.t())
    e1.record(); torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / iters
    return {"ms": ms, "tflops": 2 * M * N * K / ms / 1e9, "torch_ms": ms_t, "torch_tflops": 2 * M * N * K / ms_t / 1e9}
res["t_48000x1280x1280"] = bench(48000, 1280, 1280)
res["t_48000x3840x1280"] = bench(48000, 3840, 1280)
res["t_48000x5120x1280_gelu"] = bench(48000, 5120, 1280, gelu=True)
res["t_48000x1280x5120"] = bench(48000, 1280, 5120)
res["t_8192^3"] = bench(8192, 8192, 8192, iters=5)
print(json.dumps(res, indent=1))
