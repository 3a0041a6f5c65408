"""Dev tool: throughput of the frame-writing CLI (decode -> packed batches -> encode -> pinned D2H ring -> .pt writers)
on a synthetic corpus of PCM16 WAV files. The model load is measured by a second run with --skip_existing (nothing left
to do) and subtracted.   python tools/bench_cli.py [n_files] [ssl_type]"""
import os, sys, time, tempfile, shutil
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from interspeech_ser_b200 import audio_io
from interspeech_ser_b200.cli import main_speech

n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ssl = sys.argv[2] if len(sys.argv) > 2 else "microsoft/wavlm-large"
root = tempfile.mkdtemp(prefix="serenc_cli_")
wav_dir, out_dir = os.path.join(root, "wav"), os.path.join(root, "feat")
os.makedirs(wav_dir)
rng = np.random.default_rng(7)
secs = 0.0
for i in range(n_files):
    n = int(rng.integers(2 * 16000, 12 * 16000))
    secs += n / 16000.0
    audio_io.write_wav(os.path.join(wav_dir, f"utt_{i:06d}.wav"), (rng.standard_normal(n) * 0.0886).astype(np.float32))
argv = ["--ssl_type", ssl, "--wav_dir", wav_dir, "--save_path", out_dir, "--random_init", "--use_average", "y", "--num_workers", "8"]
t0 = time.time(); rc = main_speech(argv); t1 = time.time()
assert rc == 0 and len(os.listdir(out_dir)) == n_files
t2 = time.time(); main_speech(argv + ["--skip_existing"]); t3 = time.time()
size = sum(os.path.getsize(os.path.join(out_dir, f)) for f in os.listdir(out_dir))
work = (t1 - t0) - (t3 - t2)
print(f"RESULT files={n_files} audio_s={secs:.0f} total={t1 - t0:.1f}s load_only={t3 - t2:.1f}s -> processing {work:.2f}s = "
      f"{secs / work:.0f} audio-s/s, {n_files / work:.0f} files/s, {size / work / 1e9:.2f} GB/s of .pt written ({size / 1e9:.2f} GB)")
shutil.rmtree(root)
