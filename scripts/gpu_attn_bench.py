import ctypes as C, sys, torch
sys.path.insert(0, ".")
from manual_whisper_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0")
B, T, H = 32, 1500, 20; d = H * 64
H16 = _lib.storage_dtype(); qkv = torch.randn(B * T, 3 * d, device=dev).to(H16); out = torch.empty(B * T, d, device=dev, dtype=H16)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3): lib.mw_attention_h16(qkv.data_ptr(), out.data_ptr(), B, T, H, st)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): lib.mw_attention_h16(qkv.data_ptr(), out.data_ptr(), B, T, H, st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fl = 4.0 * T * T * d * B
print(f"attention B={B}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s")
q, k, v = [t.view(B, T, H, 64).transpose(1, 2) for t in qkv.split(d, dim=1)]
for _ in range(3): torch.nn.functional.scaled_dot_product_attention(q, k, v)
e0.record()
for _ in range(10): torch.nn.functional.scaled_dot_product_attention(q, k, v)
e1.record(); torch.cuda.synchronize()
ms2 = e0.elapsed_time(e1) / 10
print(f"torch sdpa: {ms2*1e3:.1f} us  {fl/ms2/1e9:.1f} TFLOP/s")
