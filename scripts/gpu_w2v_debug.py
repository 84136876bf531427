import sys, os, ctypes as C
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
from dataclasses import replace
from manual_whisper_b200.w2v import W2vDims, random_init_w2v
from manual_whisper_b200.alignment import AlignEngine
from manual_whisper_b200 import _lib
from oracle.wav2vec2 import OracleWav2Vec2
import torch.nn.functional as F
dims = replace(W2vDims(name="b", n_layers=2, d_model=384, n_heads=6, ffn=512, vocab=40, conv_dim=128, pos_kernel=16, pos_groups=8), feat_norm="group", stable_layer_norm=False, conv_bias=False)
sd = random_init_w2v(dims, seed=13)
eng = AlignEngine(dims, sd, device_index=0, max_batch=4, max_samples=40000)
g = np.random.default_rng(3); n=16000
audio = (0.1*g.standard_normal(n)).astype(np.float32)
em, frames = eng.emissions(torch.from_numpy(audio).cuda(), np.array([0],dtype=np.int64), np.array([n],dtype=np.int32))
T=int(frames[0]); lib=_lib.load(); st=C.c_void_p(torch.cuda.current_stream().cuda_stream)
orc = OracleWav2Vec2(dims, sd)
with torch.no_grad():
    w = torch.from_numpy(audio)
    h = F.conv1d(w.view(1,1,-1), orc.sd["wav2vec2.feature_extractor.conv_layers.0.conv.weight"], None, stride=5)
    mean = h.mean(dim=2)[0]; var = h.var(dim=2, unbiased=False)[0]
    feat = orc.features(w)
    hid = orc.hidden(w)
stats = torch.empty(1,128,2,device="cuda"); lib.mw_w2v_debug_copy(eng.handle, 2, stats.data_ptr(), stats.numel()*4, st)
torch.cuda.synchronize(); stats=stats.cpu()
print("gn mean err", (stats[0,:,0]-mean).abs().max().item(), "rstd err", (stats[0,:,1]-1/torch.sqrt(var+1e-5)).abs().max().item(), "rstd typical", (1/torch.sqrt(var+1e-5)).mean().item())
T0 = h.shape[2]
c0 = torch.empty(1,T0,128,device="cuda",dtype=_lib.storage_dtype()); lib.mw_w2v_debug_copy(eng.handle, 3, c0.data_ptr(), c0.numel()*2, st)
ref0 = F.gelu(F.group_norm(h, 128, orc.sd["wav2vec2.feature_extractor.conv_layers.0.layer_norm.weight"], orc.sd["wav2vec2.feature_extractor.conv_layers.0.layer_norm.bias"], 1e-5))[0].transpose(0,1)
torch.cuda.synchronize(); print("conv0 err", (c0[0].float().cpu()-ref0).abs().max().item(), ref0.abs().max().item())
f = torch.empty(T,128,device="cuda"); lib.mw_w2v_debug_copy(eng.handle, 0, f.data_ptr(), f.numel()*4, st); torch.cuda.synchronize()
print("feat err", (f.cpu()-feat).abs().max().item(), feat.abs().max().item())
x = torch.empty(T,384,device="cuda"); lib.mw_w2v_debug_copy(eng.handle, 1, x.data_ptr(), x.numel()*4, st); torch.cuda.synchronize()
print("hidden err", (x.cpu()-hid).abs().max().item(), hid.abs().max().item())
if os.environ.get("MW_W2V_STOP"):
    with torch.no_grad():
        d = dims
        xo = orc._lin(orc._ln(feat, "wav2vec2.feature_projection.layer_norm"), "wav2vec2.feature_projection.projection")
        xp = xo.transpose(0, 1).unsqueeze(0)
        pos = F.conv1d(xp, orc.w_pos, orc.sd["wav2vec2.encoder.pos_conv_embed.conv.bias"], padding=d.pos_kernel // 2, groups=d.pos_groups)[:, :, :-1]
        x2 = xo + F.gelu(pos)[0].transpose(0, 1)
    print("proj err", (x.cpu()-xo).abs().max().item(), "proj+pos err", (x.cpu()-x2).abs().max().item(), x2.abs().max().item())
    err = (x.cpu()-x2).abs()
    print("per-group max err", [round(err[:, g*48:(g+1)*48].max().item(),3) for g in range(8)], "per-col first group", [round(v,3) for v in err[:, :48].max(0).values.tolist()[:12]])
