"""Scratch GPU check: attention / layernorm kernels vs torch, encoder vs oracle, large-v3 encoder timing."""
import ctypes as C, json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from manual_whisper_b200 import _lib
from manual_whisper_b200.config import model_dims, custom_dims
from manual_whisper_b200.weights import random_init
from manual_whisper_b200.engine import Engine
from oracle.model import OracleWhisper
lib = _lib.load(); dev = torch.device("cuda:0"); torch.manual_seed(0)
res = {}
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
# layernorm
for d in (384, 1280):
    x = torch.randn(1000, d, device=dev) * 3 + 1; g = torch.randn(d, device=dev); b = torch.randn(d, device=dev)
    out = torch.empty(1000, d, device=dev, dtype=torch.bfloat16)
    _lib.check(lib.mw_layernorm(x.data_ptr(), g.data_ptr(), b.data_ptr(), out.data_ptr(), 1000, d, st()), "ln")
    ref = torch.nn.functional.layer_norm(x, (d,), g, b, 1e-5)
    res[f"ln_{d}"] = (out.float() - ref).abs().max().item()
# attention
for (B, T, H) in [(1, 128, 1), (1, 200, 2), (2, 1500, 6), (1, 1500, 20)]:
    d = H * 64
    qkv = (torch.randn(B * T, 3 * d, device=dev)).bfloat16()
    out = torch.zeros(B * T, d, device=dev, dtype=torch.bfloat16)
    _lib.check(lib.mw_attention_bf16(qkv.data_ptr(), out.data_ptr(), B, T, H, st()), "att")
    torch.cuda.synchronize()
    q, k, v = [t.float().view(B, T, H, 64).transpose(1, 2) for t in qkv.split(d, dim=1)]
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, d)
    res[f"att_{B}x{T}x{H}"] = [(out.float() - ref).abs().max().item(), ref.abs().max().item()]
print(json.dumps(res, indent=1), flush=True)
# encoder vs oracle
def enc_check(dims, scheme, B):
    sd = random_init(dims, scheme=scheme)
    eng = Engine(dims, sd, 0, max_batch=B)
    mel = (torch.randn(B, dims.n_mels, 2 * dims.n_audio_ctx) * 0.5).clamp(-1.5, 1.5)
    got = eng.encode(mel.to(dev)).float().cpu()
    with torch.no_grad():
        ref = OracleWhisper(dims, sd).encode(mel)
        emu = OracleWhisper(dims, sd, emulate_bf16=True).encode(mel)
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    return {"rel_l2_vs_fp32": rel(got, ref), "rel_l2_vs_emu": rel(got, emu), "maxabs_vs_fp32": (got - ref).abs().max().item(),
            "maxabs_vs_emu": (got - emu).abs().max().item(), "emu_vs_fp32": rel(emu, ref), "ref_absmax": ref.abs().max().item()}
small = custom_dims("test-small", 80, 128, 2, 2, 2, 512, 2048, n_audio_ctx=200)
for scheme in ("survey", "lively"):
    res[f"enc_small_{scheme}"] = enc_check(small, scheme, 3)
    res[f"enc_tiny_{scheme}"] = enc_check(model_dims("tiny"), scheme, 2)
print(json.dumps(res, indent=1), flush=True)
# large-v3 timing with device-generated weights
dims = model_dims("large-v3")
from manual_whisper_b200.weights import _keys
g = torch.Generator(device=dev); g.manual_seed(1)
sd = {}
for name, shape, kind in _keys(dims):
    if kind == "g": sd[name] = torch.ones(shape, device=dev)
    elif kind == "beta": sd[name] = torch.zeros(shape, device=dev)
    else: sd[name] = (torch.randn(shape, device=dev, generator=g) * 0.02)
sd["model.encoder.embed_positions.weight"] = torch.zeros(1500, dims.d_model, device=dev)
B = 32
eng = Engine(dims, sd, 0, max_batch=B)
del sd
mel = torch.randn(B, 128, 3000, device=dev) * 0.5
for _ in range(2): eng.encode(mel)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): out = eng.encode(mel)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
res["large_v3_encoder_B32_ms"] = ms
res["large_v3_encoder_tflops"] = 2.2738 * B / ms * 1e3 / 1e3
res["finite"] = bool(torch.isfinite(out.float()).all())
res["workspace_GB"] = eng.workspace_bytes / 1e9
print(json.dumps(res, indent=1))
