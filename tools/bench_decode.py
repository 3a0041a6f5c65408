"""Dev tool (CPU): host WAV ingest strategies on a synthetic PCM16 corpus - inline decode, N full-decode threads, and
the CLI's arrangement (reader threads in chunks of 64 + one decoder thread).   python tools/bench_decode.py [n_files]"""
import os, sys, tempfile, time, shutil
from concurrent.futures import ThreadPoolExecutor
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from interspeech_ser_b200 import audio_io

n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 800
root = tempfile.mkdtemp(prefix="serenc_dec_")
rng = np.random.default_rng(0)
paths, secs = [], 0.0
for i in range(n_files):
    n = int(rng.integers(2 * 16000, 12 * 16000))
    secs += n / 16000.0
    p = os.path.join(root, f"{i}.wav")
    audio_io.write_wav(p, (rng.standard_normal(n) * 0.0886).astype(np.float32))
    paths.append(p)

def report(name, t):
    print(f"{name:48s} {t / n_files * 1e3:7.3f} ms/file  {secs / t:9.0f} audio-s/s")

t0 = time.time(); [audio_io.load_audio(p) for p in paths]; report("inline, one thread", time.time() - t0)
for nt in (2, 4, 8):
    t0 = time.time()
    with ThreadPoolExecutor(nt) as ex:
        list(ex.map(audio_io.load_audio, paths))
    report(f"{nt} threads, full decode per file", time.time() - t0)

def decode_chunk(ps):
    return [audio_io.load_audio(p) for p in ps]

for nt in (2, 4):
    chunks = [paths[i:i + 16] for i in range(0, n_files, 16)]
    t0 = time.time()
    with ThreadPoolExecutor(nt) as ex:
        [y for c in ex.map(decode_chunk, chunks) for y in c]
    report(f"{nt} threads, chunks of 16 files (the CLI)", time.time() - t0)

def read_chunk(ps):
    return [audio_io.read_bytes(p) for p in ps]

for nt in (4, 8):
    chunks = [paths[i:i + 64] for i in range(0, n_files, 64)]
    t0 = time.time()
    with ThreadPoolExecutor(nt) as readers, ThreadPoolExecutor(1) as dec:
        def work():
            return [audio_io.load_audio_bytes(d, p) for c, ds in zip(chunks, readers.map(read_chunk, chunks)) for p, d in zip(c, ds)]
        dec.submit(work).result()
    report(f"{nt} reader threads (chunks of 64) + 1 decoder", time.time() - t0)
shutil.rmtree(root)
