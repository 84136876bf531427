"""Audio decode on the GPU: mw_pcm_resample throughput on 1 hour of PCM already in HBM, and from
pinned host memory; beside the numpy restatement (oracle/) on a 60-s sample.  Prints one JSON object."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import manual_whisper_b200 as mw
rows = []
for rate, ch in ((48000, 2), (44100, 2), (48000, 1)):
    n = rate * 3600
    pcm = torch.randint(-20000, 20000, (n * ch,), dtype=torch.int16)
    pinned = pcm.pin_memory()
    d = pinned.cuda()
    for _ in range(2):
        out = mw.decode_pcm_device(d, ch, rate)
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = mw.decode_pcm_device(d, ch, rate); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    t0 = time.perf_counter(); out = mw.decode_pcm_device(pinned, ch, rate); torch.cuda.synchronize(); e2e = time.perf_counter() - t0
    nbytes = n * ch * 2 + out.numel() * 4
    row = {"rate": rate, "channels": ch, "hours": 1, "kernel_ms": ms, "GBps": nbytes / ms / 1e6, "rtfx_resident": 3600 / (ms / 1e3),
           "from_pinned_host_s": e2e, "rtfx_from_host": 3600 / e2e}
    if "--no-cpu" not in sys.argv:
        from oracle.resample import decode_pcm16
        samp = pcm[: rate * 60 * ch].numpy()
        t0 = time.perf_counter(); decode_pcm16(samp, ch, rate); dt = time.perf_counter() - t0
        row["cpu_numpy_rtfx_60s_sample"] = 60 / dt
    rows.append(row)
    del d, out, pinned, pcm
print(json.dumps(rows, indent=1))
