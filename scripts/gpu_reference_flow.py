"""The reference's transcribe_audio() steps 1-2 (/root/reference/transcribe.py:107-131) on one B200 for a 1-hour synthetic
recording delivered as a 48 kHz stereo WAV: decode -> VAD windows -> large-v3 ASR -> wav2vec2 forced alignment.
Weights are random-init (no checkpoints offline).  Prints one JSON object with the wall time of every stage."""
import json, os, sys, time, wave, tempfile, warnings
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import manual_whisper_b200 as mw
from manual_whisper_b200.config import model_dims
from bench import device_weights, MODEL, BATCH
warnings.simplefilter("ignore")
hours = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
dev = torch.device("cuda:0")
audio16, turns = mw.synthetic_speech(3600.0 * hours, seed=1)
# a 48 kHz stereo rendition of the same recording (linear upsampling is enough for a timing run)
up = np.repeat(audio16, 3)
pcm = np.clip(np.rint(np.stack([up, up], 1) * 32768), -32768, 32767).astype(np.int16)
path = os.path.join(tempfile.mkdtemp(), "meeting.wav")
with wave.open(path, "wb") as w:
    w.setnchannels(2); w.setsampwidth(2); w.setframerate(48000); w.writeframes(pcm.tobytes())
del up, pcm
pipe = mw.load_model(MODEL, "cuda", compute_type="float16", language="zh", asr_options={"beam_size": 1},
                     vad_model=mw.InjectedVad(turns), model=device_weights(model_dims(MODEL), dev, seed=1234), max_batch=BATCH,
                     streams_per_device=8)
model_a, meta = mw.load_align_model("zh", "cuda", max_batch=16)
res = {"audio_s": 3600.0 * hours, "wav_MB": os.path.getsize(path) / 1e6}
for it in range(2):             # first pass warms graphs and caches
    torch.cuda.synchronize(); t0 = time.perf_counter()
    audio = mw.load_audio_device(path)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    result = pipe.transcribe(audio, batch_size=BATCH, language="zh")
    torch.cuda.synchronize(); t2 = time.perf_counter()
    # no vocabulary offline: give every segment 5 characters per second so the CTC work is realistic
    segs = [dict(s, text="".join("abcdefghij"[(i + k) % 10] for k in range(int((s["end"] - s["start"]) * 5))))
            for i, s in enumerate(result["segments"])]
    aligned = mw.align(segs, model_a, meta, audio, "cuda", return_char_alignments=False)
    torch.cuda.synchronize(); t3 = time.perf_counter()
res.update(windows=len(result["segments"]), words=len(aligned["word_segments"]), read_and_decode_s=t1 - t0, transcribe_s=t2 - t1,
           align_s=t3 - t2, total_s=t3 - t0, rtfx_total=res["audio_s"] / (t3 - t0))
print(json.dumps(res, indent=1))
