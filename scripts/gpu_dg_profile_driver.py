"""Launches the decode GEMM at large-v3 shapes a few times each (driver for ncu duration / full-set captures)."""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from manual_whisper_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0"); H = _lib.storage_dtype()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
shapes = {"dxd": (1280, 1280), "qkv": (3840, 1280), "fc1": (5120, 1280), "fc2": (1280, 5120), "logits": (51866, 1280)}
want = os.environ.get("DG_SHAPES", ",".join(shapes)).split(",")
rows = [int(r) for r in os.environ.get("DG_ROWS", "32,128,256").split(",")]
for name in want:
    N, K = shapes[name]
    copies = 3
    w = (torch.randn(copies, N, K, device=dev) * 0.02).to(H)
    bias = torch.randn(N, device=dev)
    for R in rows:
        x = (torch.randn(R, K, device=dev) * 0.5).to(H)
        o = torch.empty(R, N, device=dev, dtype=H)
        for i in range(3):
            _lib.check(lib.mw_decode_gemm_h16(x.data_ptr(), w[i % copies].data_ptr(), bias.data_ptr(), None, o.data_ptr(), R, N, K, 0, st), "dg")
torch.cuda.synchronize()
print("ok")
