"""Dev tool: throughput of the frame-writing CLI (decode -> packed batches -> encode -> pinned D2H ring -> .pt writers)
on a synthetic corpus of PCM16 WAV files. The model load is measured by a second run with --skip_existing (nothing left
to do) and subtracted.

    python tools/bench_cli.py [n_files] [ssl_type]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_cli.py [n_files]

Under torchrun, rank 0 writes the corpus, every rank runs the CLI on its file shard (cli.py shards by size on disk), and
the slowest rank's time counts. SERENC_CLI_TIMING=1 prints the main-thread phase times of every rank."""
import os, sys, time, shutil
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from interspeech_ser_b200 import audio_io
from interspeech_ser_b200.cli import main_speech

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ssl = sys.argv[2] if len(sys.argv) > 2 else "microsoft/wavlm-large"
root = os.path.join("/tmp", f"serenc_cli_{os.environ.get('MASTER_PORT', os.getpid())}")
wav_dir, out_dir = os.path.join(root, "wav"), os.path.join(root, "feat")
dist = None
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")
secs = 0.0
if rank == 0:
    shutil.rmtree(root, ignore_errors=True)
    os.makedirs(wav_dir)
    rng = np.random.default_rng(7)
    for i in range(n_files):
        n = int(rng.integers(2 * 16000, 12 * 16000))
        secs += n / 16000.0
        audio_io.write_wav(os.path.join(wav_dir, f"utt_{i:06d}.wav"), (rng.standard_normal(n) * 0.0886).astype(np.float32))
if dist:
    dist.barrier()
argv = ["--ssl_type", ssl, "--wav_dir", wav_dir, "--save_path", out_dir, "--random_init", "--use_average", "y", "--num_workers", "8"]


def timed(extra):
    if dist:
        dist.barrier()
    t0 = time.time()
    rc = main_speech(argv + extra)
    dt = time.time() - t0
    assert rc == 0
    if dist:
        import torch
        t = torch.tensor([dt], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return dt


total = timed([])
load_only = timed(["--skip_existing"])
if rank == 0:
    assert len(os.listdir(out_dir)) == n_files
    size = sum(os.path.getsize(os.path.join(out_dir, f)) for f in os.listdir(out_dir))
    work = total - load_only
    print(f"RESULT ranks={world} files={n_files} audio_s={secs:.0f} total={total:.1f}s load_only={load_only:.1f}s -> processing {work:.2f}s = "
          f"{secs / work:.0f} audio-s/s, {n_files / work:.0f} files/s, {size / work / 1e9:.2f} GB/s of .pt written ({size / 1e9:.2f} GB)")
    shutil.rmtree(root)
if dist:
    dist.barrier()
    dist.destroy_process_group()
