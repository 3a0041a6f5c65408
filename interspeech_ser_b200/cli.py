"""Command-line drivers with the reference's flags and on-disk layout.

    python preprocessing/preprocess_speech.py  --ssl_type microsoft/wavlm-large --wav_dir W --save_path S [--n_layer -1] [--use_average y]
    python preprocessing/preprocess_whisper.py --ssl_type openai/whisper-large-v3 --wav_dir W --save_path S ...

Same contract as preprocessing/preprocess_speech.py:13-22,69-73 and preprocess_whisper.py of the reference: one
`<basename>.pt` per input file holding a 2-D float32 tensor [T, D]; a failing file prints
`Failed to process <path>: <error>` and the run continues; an unknown model aborts with the reference's message.
What changes is the schedule: files are decoded by `--num_workers` host threads, length-bucketed into packed
batches and (under torchrun) sharded over the GPUs of the box; tensors are saved as contiguous CPU tensors
(identical `torch.load` result, no hidden 7.7 MB storage, no CUDA pinning of the consumer).
"""
from __future__ import annotations

import argparse
import os
import time
import sys
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional

import numpy as np


def build_parser(whisper: bool) -> argparse.ArgumentParser:
    p = argparse.ArgumentParser()
    # the reference's flags (preprocess_speech.py:13-22)
    p.add_argument("--seed", type=int, default=7)
    p.add_argument("--ssl_type", type=str, default="wavlm-large")
    p.add_argument("--save_path", type=str, default="./")
    p.add_argument("--wav_dir", type=str, default="./")
    p.add_argument("--num_workers", type=int, default=4)
    p.add_argument("--n_layer", type=int, default=-1)
    p.add_argument("--use_average", type=str, default="n")
    # additions
    p.add_argument("--random_init", action="store_true", help="random weights (no checkpoint available offline)")
    p.add_argument("--frame_budget", type=int, default=28416, help="max frames per packed batch (111 x 256: whole GEMM waves)")
    p.add_argument("--skip_existing", action="store_true", help="resume: skip files whose .pt already exists")
    p.add_argument("--window_files", type=int, default=4096,
                   help="files decoded, batched and encoded at a time (the next window is decoded while this one runs); "
                        "bounds host memory on a 100k-file corpus")
    p.add_argument("--pooled_path", type=str, default="", help="also save masked-mean pooled embeddings {names, embeddings[N, D]}")
    p.add_argument("--float_upload", action="store_true",
                   help="decode 16-bit PCM to float32 on the host (default: upload the int16 samples and scale on the GPU; same result)")
    p.add_argument("--checkpoint", type=str, default="",
                   help="weights file/dir for --ssl_type's architecture; a peft LoRA classifier state dict "
                        "(preprocess_speech_pretrained.py:173) is merged into dense weights at load")
    if not whisper:
        p.add_argument("--compat_layer_from_dir_count", action="store_true",
                       help="literal preprocess_speech.py:41,67 behaviour: index hidden_states by the number of files already in --save_path")
    else:
        p.add_argument("--crop_cap_1500", action="store_true",
                       help="crop to min(ceil(len/320), 1500) frames instead of the script's literal min(.., hidden_size) (preprocess_whisper.py:75)")
    return p


def run(argv: Optional[List[str]], whisper: bool) -> int:
    args = build_parser(whisper).parse_args(argv)
    import torch

    from .audio_io import load_audio_bytes, read_bytes
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from .configs import ARCH_WHISPER
    from .modeling import AutoModel
    from . import scheduler

    average = args.use_average == "y"
    print(f"Using average = {average}")
    if not torch.cuda.is_available():
        print("Error: no CUDA device visible; this extractor has no CPU path.")
        return 2
    device = torch.device(f"cuda:{local_rank}")
    print(f"Using device = {device}")

    # Under torchrun every rank must shard the SAME work list: rank 0 alone looks at the live save_path (the count
    # behind --compat_layer_from_dir_count and the --skip_existing filter) and broadcasts what it saw, BEFORE any rank
    # has written a file. (Ranks load the model at different speeds; a rank listing the directory after another one
    # started writing would shard a different list: files encoded twice, others never.)
    os.makedirs(args.save_path, exist_ok=True)
    wav_files = sorted(os.listdir(args.wav_dir))

    def out_path(name):
        return os.path.join(args.save_path, os.path.splitext(os.path.basename(name))[0] + ".pt")

    if rank == 0:
        n_existing = len(os.listdir(args.save_path))
        todo = [w for w in wav_files if not (args.skip_existing and os.path.exists(out_path(w)))]
        shared = [n_existing, todo]
    else:
        shared = [None, None]
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("gloo")
        dist.broadcast_object_list(shared, src=0)   # doubles as the barrier that keeps every writer behind rank 0's listing
    n_existing, todo = shared
    print(f"Save path = {args.save_path} created. It has {n_existing} files in it.")

    print(f"{len(wav_files)} file are going to be processed...")
    print(f"Checking files in {args.wav_dir}")
    missing = [os.path.join(args.wav_dir, w) for w in wav_files if not os.path.isfile(os.path.join(args.wav_dir, w))]
    if missing:
        print("Missing files:")
        for m in missing:
            print(f" - {m}")
        print("Something went wrong, make sure everything is correct before running again!")
        return 1

    print(f"Extracting features using {args.ssl_type}")
    try:
        if args.checkpoint:
            model = AutoModel.from_pretrained(args.checkpoint, device=local_rank, config_name=args.ssl_type)
        else:
            model = AutoModel.from_pretrained(args.ssl_type, device=local_rank, random_init=args.random_init or None, seed=0)
    except OSError:
        print(f"Error: No pretrained model found with the name {args.ssl_type}")
        print("Something went wrong, make sure everything is correct before running again!")
        return 1
    cfg = model.cfg
    if (cfg.arch == ARCH_WHISPER) != whisper:
        print(f"Error: {args.ssl_type} is {'not ' if whisper else ''}a Whisper model; use the other script")
        return 1

    layer = args.n_layer
    if not whisper and getattr(args, "compat_layer_from_dir_count", False):
        layer = n_existing  # preprocess_speech.py:41,67

    # ---- rank sharding on FILE SIZE (a PCM WAV's size is its length): every rank decodes only its own files ----
    sizes = [float(os.path.getsize(os.path.join(args.wav_dir, w))) for w in todo]
    mine_files = [todo[i] for i in scheduler.shard_by_cost(sizes, world)[rank]]   # ascending size: windows hold similar lengths

    # ---- host ingest, one window ahead of the GPU ----
    # Measured (tools/bench_decode.py, 8-vCPU container): PCM decode is memory- and GIL-bound: inline 15 k audio-s/s,
    # 4 threads 19 k, 8 threads 7 k (they fight over the interpreter), reader threads + one decoder thread 11 k. So:
    # at most 4 decode threads however large --num_workers is, each task a chunk of 16 files.
    def decode_one(name):
        path = os.path.join(args.wav_dir, name)
        try:
            y, _ = load_audio_bytes(read_bytes(path), path, sr=16000, keep_int16=not args.float_upload)
            if not whisper and len(y) < 400:
                raise ValueError(f"{len(y)} samples is shorter than the encoder's 400-sample receptive field")
            if len(y) == 0:
                raise ValueError("empty audio")
            return name, y
        except Exception as e:  # noqa: BLE001  (reference: preprocess_speech.py:72-73)
            print(f"Failed to process {path}: {e}")
            return name, None

    def decode_chunk(names_chunk):
        return [decode_one(n) for n in names_chunk]

    win = max(1, args.window_files)
    windows = [mine_files[i:i + win] for i in range(0, len(mine_files), win)]
    readers = ThreadPoolExecutor(max_workers=max(1, min(4, args.num_workers)))
    decoder = ThreadPoolExecutor(max_workers=1)          # owns a window: the main thread only waits on one future
    writer = ThreadPoolExecutor(max_workers=max(1, args.num_workers))
    from .engine import DownloadRing
    downloads = DownloadRing(device)

    def decode_window(names_w):
        chunks = [names_w[i:i + 16] for i in range(0, len(names_w), 16)]
        return [r for c in readers.map(decode_chunk, chunks) for r in c]

    def submit_window(k):
        return decoder.submit(decode_window, windows[k]) if k < len(windows) else None

    pooled_rows = {}
    pooled_pending = []
    futures = []
    n_done = 0
    timing = {"decode_wait": 0.0, "encode_call": 0.0, "download_call": 0.0, "writer_drain": 0.0} if os.environ.get("SERENC_CLI_TIMING") else None
    pending = submit_window(0)
    for k in range(len(windows)):
        tt = time.time()
        loaded = pending.result()
        if timing is not None:
            timing["decode_wait"] += time.time() - tt
        pending = submit_window(k + 1)           # decoded while this window is on the GPU
        names = [n for n, y in loaded if y is not None]
        waves = [y for _, y in loaded if y is not None]
        if not waves:
            continue
        for bt in scheduler.make_batches(cfg, [len(y) for y in waves], frame_budget=args.frame_budget):
            idx = bt.indices
            try:
                if layer >= cfg.num_hidden_layers + 1 or layer < -(cfg.num_hidden_layers + 1):
                    raise IndexError("tuple index out of range")  # what hidden_states[N] raises in the reference
                kw = dict(layer=layer, average=average, want_frames=True, want_pooled=bool(args.pooled_path))
                if whisper:
                    kw["literal_crop"] = not args.crop_cap_1500
                tt = time.time()
                res = model.extract([waves[i] for i in idx], **kw)
                if timing is not None:
                    timing["encode_call"] += time.time() - tt
                    tt = time.time()
                # one D2H of the packed frames on the copy stream into a pinned ring; a writer thread waits for it, clones
                # each utterance's rows (contiguous tensors that own their storage: no 1500-frame storage behind a view,
                # defect D3) and saves them, while this thread goes on to the next batch
                host, done, slot = downloads.download(res.packed)
                if timing is not None:
                    timing["download_call"] += time.time() - tt
                pooled_dev = res.pooled
                paths = [out_path(names[i]) for i in idx]

                def write_batch(host=host, done=done, slot=slot, ranges=res.ranges, paths=paths, keep=res.packed):
                    try:
                        done.synchronize()
                        for (a, e), path in zip(ranges, paths):
                            try:
                                torch.save(host[a:e].clone(), path)
                            except Exception as ex:  # noqa: BLE001  (reference: error-and-continue per file)
                                print(f"Failed to process {path}: {ex}")
                    finally:
                        downloads.release(slot)
                futures.append(writer.submit(write_batch))
                if pooled_dev is not None:   # pooled rows come back asynchronously too; resolved after the last batch
                    pc = torch.empty(pooled_dev.shape, dtype=torch.float32).pin_memory()
                    pc.copy_(pooled_dev, non_blocking=True)
                    pooled_pending.append((pc, [names[i] for i in idx]))
                n_done += len(idx)
            except Exception as e:  # noqa: BLE001
                for i in idx:
                    print(f"Failed to process {os.path.join(args.wav_dir, names[i])}: {e}")
    tt = time.time()
    for f in futures:        # writers drain (the download ring's slots are their back-pressure on the loop above)
        f.result()
    decoder.shutdown()
    readers.shutdown()
    writer.shutdown()
    torch.cuda.synchronize(device)
    if timing is not None:
        timing["writer_drain"] = time.time() - tt
        print("host time per phase [s] (main thread): " + ", ".join(f"{k} {v:.2f}" for k, v in timing.items()))
    for pc, ns in pooled_pending:
        for j, n in enumerate(ns):
            pooled_rows[n] = pc[j].clone()

    if args.pooled_path:
        rows = pooled_rows
        if world > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                dist.init_process_group("gloo")
            gathered = [None] * world if rank == 0 else None
            dist.gather_object(rows, gathered, dst=0)  # the path's only exchange: final host gather
            if rank == 0:
                rows = {}
                for g in gathered:
                    rows.update(g)
        if rank == 0:
            keys = sorted(rows)
            emb = torch.stack([rows[k] for k in keys]) if keys else torch.empty(0, cfg.hidden_size)
            torch.save({"names": keys, "embeddings": emb}, args.pooled_path)
    print(f"Done: {n_done} utterances on rank {rank}/{world}.")
    return 0


def main_speech(argv: Optional[List[str]] = None) -> int:
    return run(argv, whisper=False)


def main_whisper(argv: Optional[List[str]] = None) -> int:
    return run(argv, whisper=True)


# --------------------------------------------------------------------------------------------------
# text branch: preprocessing/preprocess_roberta.py
# --------------------------------------------------------------------------------------------------
def build_roberta_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser()
    # the reference's flags (preprocess_roberta.py:13-20)
    p.add_argument("--seed", type=int, default=7)
    p.add_argument("--roberta_type", type=str, default="roberta")
    p.add_argument("--df_path", type=str, default="./")
    p.add_argument("--save_path", type=str, default="./")
    p.add_argument("--num_workers", type=int, default=4)
    p.add_argument("--max_len", type=int, default=80)
    p.add_argument("--use_average", type=str, default="n")
    # additions
    p.add_argument("--random_init", action="store_true", help="random weights (no checkpoint available offline)")
    p.add_argument("--tokenizer_path", type=str, default="", help="directory with vocab.json + merges.txt (default: --roberta_type)")
    p.add_argument("--batch_texts", type=int, default=256, help="texts per encode call")
    p.add_argument("--skip_existing", action="store_true")
    return p


def main_roberta(argv: Optional[List[str]] = None) -> int:
    """preprocess_roberta.py: CSV (columns `transcription`, `FileName`) -> <basename>.pt holding [max_len, D] fp32 — every
    position of the max_length-padded sequence, pad positions included, as the reference saves them (:49-74)."""
    args = build_roberta_parser().parse_args(argv)
    import pandas as pd
    import torch

    from .configs import ARCH_TEXT
    from .modeling import AutoModel
    from .text import RobertaTokenizer

    average = args.use_average == "y"
    print(f"Using average = {average}")
    if not torch.cuda.is_available():
        print("Error: no CUDA device visible; this extractor has no CPU path.")
        return 2
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device(f"cuda:{local_rank}")
    print(f"Using device = {device}")
    os.makedirs(args.save_path, exist_ok=True)
    print(f"Save path = {args.save_path} created. It has {len(os.listdir(args.save_path))} files in it.")
    print(f"Reading dataframe {args.df_path}")
    try:
        df = pd.read_csv(args.df_path)
        texts = [str(t) for t in df.transcription.values]
        names = [str(n) for n in df.FileName.values]
    except Exception as e:  # noqa: BLE001  (preprocess_roberta.py:91-95)
        print(f"Error reading dataframe from {args.df_path}: {e}")
        print("Something went wrong, make sure everything is correct before running again!")
        return 1
    print(f"Extracting features using {args.roberta_type}")
    try:
        tokenizer = RobertaTokenizer.from_pretrained(args.tokenizer_path or args.roberta_type)
        model = AutoModel.from_pretrained(args.roberta_type, device=local_rank, random_init=args.random_init or None, seed=0)
        if model.cfg.arch != ARCH_TEXT:
            raise OSError(f"{args.roberta_type} is not a text encoder")
    except OSError:
        print(f"Error: No pretrained model found with the name {args.roberta_type}")
        print("Something went wrong, make sure everything is correct before running again!")
        return 1

    def out_path(name):
        return os.path.join(args.save_path, os.path.splitext(os.path.basename(name))[0] + ".pt")

    todo = [i for i in range(len(texts)) if i % world == rank and not (args.skip_existing and os.path.exists(out_path(names[i])))]
    writer = ThreadPoolExecutor(max_workers=max(1, args.num_workers))
    futures = []
    n_done = 0
    for k in range(0, len(todo), max(1, args.batch_texts)):
        idx = todo[k:k + max(1, args.batch_texts)]
        try:
            enc = tokenizer([texts[i] for i in idx], padding="max_length", truncation=True, max_length=args.max_len, return_tensors="pt")
            res = model.extract_tokens(enc["input_ids"], enc["attention_mask"], layer=-1, average=average, want_frames=True)
            host = res.packed.cpu()      # one D2H for the batch
            T = args.max_len

            def write(host=host, idx=idx, T=T):
                for j, i in enumerate(idx):
                    try:
                        torch.save(host[j * T:(j + 1) * T].clone(), out_path(names[i]))
                    except Exception as ex:  # noqa: BLE001
                        print(f"Failed to process {names[i]}: {ex}")
            futures.append(writer.submit(write))
            n_done += len(idx)
        except Exception as e:  # noqa: BLE001  (reference: error-and-continue, :75-76)
            for i in idx:
                print(f"Failed to process {names[i]}: {e}")
    for f in futures:
        f.result()
    writer.shutdown()
    print(f"Done: {n_done} texts on rank {rank}/{world}.")
    return 0
