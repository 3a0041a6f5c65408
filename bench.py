#!/usr/bin/env python
"""Benchmark of the embedding-extraction hot path: audio-seconds encoded per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--workloads all|none|a,b,c]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
        (what the driver runs; a bare `python bench.py --gpus N` with N > 1 re-executes itself as that launch)
    python bench.py --impl reference ...     # the reference's own CPU implementation (HF transformers) on host cores

One "step" = one pass of the hot path over one batch of synthetic utterances per GPU:
waveforms -> (normalise ->) encoder -> mean of the last four hidden states (preprocess_speech.py:56-63,
`--use_average y`) -> masked-mean pooled embedding per utterance.

  value   device-resident inputs, device-timed (CUDA events), L2 flushed between steps (untimed)
  e2e     pinned host waveforms -> H2D -> encode -> pooled [B, d] -> D2H, through the public Python API
  workloads     the other BASELINE.json configs (Whisper-large-v3 32 x 30 s, HuBERT-xlarge ragged 2-12 s, XLS-R-2b 64 x 8 s,
                WavLM-large 8 x 4 s, WavLM-large corpus sweep), each with value / e2e / roofline / kernel_breakdown,
                timed in the same run with the same method
  parity_check  after the timed loops: rows of the benched batch against the same utterance encoded alone (bit-exact),
                and the committed HuggingFace golden utterances through the benched model (cosine >= 0.999, max-rel
                <= 2e-2); the run fails on a violation
Weak scaling: every rank encodes its own shard; no data-path collective, one final gather. `--workload wavlm-large-corpus`
is the strong-scaling measurement: ONE fixed corpus sharded over the ranks by the scheduler.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WAVE_STD = 0.0886  # MSP-Podcast corpus std (reference: benchmark/model/cat_ser/7/train_norm_stat.pkl)
COS_MIN, REL_MAX = 0.999, 2e-2   # north_star tolerances (bf16 path)

# name -> model, utterance seconds (None: ragged U[lo, hi] s, scheduler-built packed batches), utterances per GPU and step
# 142 x 199 frames = 28 258 rows = 111 row-tiles of 256: every transformer GEMM is then a whole number of
# waves over the 74 CTA pairs (111 x {4, 12, 16} n-tiles = {6, 18, 24} x 74) — the scheduler's frame budget.
WORKLOADS = {
    "wavlm-large": dict(model="microsoft/wavlm-large", secs=4.0, batch=142,
                        desc="WavLM-large (random-init) embedding extraction, 4 s synthetic 16 kHz utterances "
                             "(BASELINE configs[0] utterance shape), batch 142 per GPU"),
    "wavlm-large-c1": dict(model="microsoft/wavlm-large", secs=4.0, batch=8,
                           desc="WavLM-large (random-init) embedding extraction, batch 8 x 4 s synthetic 16 kHz utterances (BASELINE configs[0]); "
                                "the public call replays a cached CUDA graph at this size"),
    "whisper-large-v3": dict(model="openai/whisper-large-v3", secs=30.0, batch=32,
                             desc="Whisper-large-v3 encoder: 128-bin log-mel frontend + encoder over 30 s synthetic audio, batch 32 per GPU (BASELINE configs[1])"),
    "hubert-xlarge": dict(model="facebook/hubert-xlarge-ls960-ft", secs=None, lo=2.0, hi=12.0, batch=256,
                          desc="HuBERT-xlarge-ls960 embedding extraction, 256 utterances per GPU of variable length U[2 s, 12 s], "
                               "length-bucketed packed batches from the scheduler (BASELINE configs[2])"),
    "hubert-xlarge-64x8s": dict(model="facebook/hubert-xlarge-ls960-ft", secs=8.0, batch=64,
                                desc="HuBERT-xlarge-ls960 embedding extraction, 8 s utterances, batch 64 per GPU"),
    "xls-r-2b": dict(model="facebook/wav2vec2-xls-r-2b", secs=8.0, batch=64,
                     desc="wav2vec2-xls-r-2b embedding extraction, batch 64 x 8 s per GPU (BASELINE configs[3])"),
    "wavlm-large-sweep": dict(model="microsoft/wavlm-large", secs=None, lo=2.0, hi=20.0, batch=512,
                              desc="WavLM-large (random-init) corpus sweep: 512 synthetic utterances per GPU, lengths U[2 s, 20 s], "
                                   "length-sorted packed batches from the scheduler (BASELINE configs[4] style)"),
    "wavlm-large-corpus": dict(model="microsoft/wavlm-large", secs=None, lo=2.0, hi=20.0, batch=4096, strong=True,
                               desc="WavLM-large (random-init), ONE fixed synthetic corpus of 4096 utterances U[2 s, 20 s] (seed 7) for the whole job: "
                                    "scheduler.plan deals its packed batches to the ranks (LPT on the FLOP model), every rank encodes its share from "
                                    "pinned host memory, rank 0 gathers and un-permutes the pooled matrix (strong scaling)"),
}
DEFAULT_EXTRAS = ["whisper-large-v3", "hubert-xlarge", "xls-r-2b", "wavlm-large-c1", "wavlm-large-sweep"]
GOLDEN = {"microsoft/wavlm-large": "microsoft__wavlm-large.npz", "openai/whisper-large-v3": "openai__whisper-large-v3.npz",
          "facebook/hubert-xlarge-ls960-ft": "facebook__hubert-xlarge-ls960-ft.npz", "facebook/wav2vec2-xls-r-2b": "facebook__wav2vec2-xls-r-2b.npz"}


def synth_wave(seed: int, n: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n, dtype=np.float32) * np.float32(WAVE_STD)).astype(np.float32)


def synth_batch(seed: int, batch: int, n_samples: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((batch, n_samples), dtype=np.float32) * np.float32(WAVE_STD)).astype(np.float32)


def metric_name(workload: str) -> str:
    model = WORKLOADS[workload]["model"]
    return "audio-seconds/sec encoded" + (" (WavLM-large)" if "wavlm" in workload else f" ({model})")


def workload_config(workload: str, world: int, batch: int | None = None, graph: bool = False) -> dict:
    """The `config` object of the JSON line; both arms (ours and --impl reference) print the same one."""
    wl = WORKLOADS[workload]
    b = batch or wl["batch"]
    strong = bool(wl.get("strong"))
    return {"workload": wl["desc"], "model": wl["model"],
            "per_gpu_batch": (b if not strong else None), "global_batch": (b * world if not strong else b),
            "utterance_seconds": wl["secs"] if wl["secs"] is not None else f"U[{wl['lo']:g}, {wl['hi']:g}]",
            "output": "mean of last 4 hidden states -> masked-mean pooled [B, d] fp32",
            "weights": "random init (seed 0)", "l2": "256 MiB buffer zeroed between timed steps (untimed)",
            "parallelism": f"utterance-sharded replicas x{world}, no data-path collective",
            "cuda_graph": bool(graph)}


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)), "samples": len(sm),
                "reasons": sorted(reasons)}


def ncu_traffic(workload: str):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/ncu_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum averaged over the four GEMMs of one WavLM-large encoder layer).
    A STATIC, profiler-side number read from a committed file, not measured by this run: null for workloads that were
    not captured."""
    if workload != "wavlm-large":
        return None
    try:
        with open(os.path.join(REPO, "profiles", "ncu_traffic.json")) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        return None


# --------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: HF transformers fp32 on the host cores (what the reference scripts call)
# --------------------------------------------------------------------------------------------------
def cpu_reference_run(workload: str, steps: int, warmup: int, n_utts: int, max_seconds: float = 150.0, padded_too: bool = False):
    """One utterance per forward, as preprocess_speech.py:45-73 / preprocess_whisper.py:45-82 do (mode i of BASELINE.md §3);
    padded_too adds mode ii, the same utterances as ONE padded batch with attention_mask."""
    import torch

    wl = WORKLOADS[workload]
    from interspeech_ser_b200 import configs
    from interspeech_ser_b200.weights import random_init

    cfg = configs.get_config(wl["model"])
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if wl["secs"] is not None:
        lens = [int(wl["secs"] * 16000)] * n_utts
    else:   # ragged workloads: a fixed-seed sample of the length distribution
        lens = [int(v) for v in np.random.default_rng(7).integers(int(wl["lo"] * 16000), int(wl["hi"] * 16000) + 1, size=n_utts)]
    waves = [synth_wave(7 + j, n) for j, n in enumerate(lens)]
    audio_per_step = sum(lens) / 16000.0
    kind = "reference"
    batch_fn = None
    try:
        import transformers as tr
        from oracle.make_golden import hf_model
        w = random_init(cfg, 0)
        model = hf_model(cfg, w)
        if cfg.family == "whisper":
            fe = tr.WhisperFeatureExtractor(feature_size=cfg.num_mel_bins)

            def one(x):
                feats = fe(x, sampling_rate=16000, return_tensors="pt")["input_features"]
                with torch.no_grad():
                    hs = model(feats, output_hidden_states=True).hidden_states
                f = torch.mean(torch.stack(hs[-4:]), dim=0).squeeze(0)
                return f[: min(int(np.ceil(len(x) / 320)), f.shape[1])].mean(0)

            def batch_fn(xs):
                feats = fe(xs, sampling_rate=16000, return_tensors="pt")["input_features"]
                with torch.no_grad():
                    hs = model(feats, output_hidden_states=True).hidden_states
                return torch.mean(torch.stack(hs[-4:]), dim=0).mean(1)
        else:
            fe = tr.Wav2Vec2FeatureExtractor(feature_size=1, sampling_rate=16000, padding_value=0.0, do_normalize=True, return_attention_mask=True)

            def one(x):
                inputs = fe(x, sampling_rate=16000, return_tensors="pt", padding=True)
                with torch.no_grad():
                    hs = model(**inputs, output_hidden_states=True).hidden_states
                return torch.mean(torch.stack(hs[-4:]), dim=0).squeeze(0).mean(0)

            def batch_fn(xs):
                inputs = fe(xs, sampling_rate=16000, return_tensors="pt", padding=True)
                with torch.no_grad():
                    hs = model(**inputs, output_hidden_states=True).hidden_states
                return torch.mean(torch.stack(hs[-4:]), dim=0).mean(1)
    except Exception:  # transformers missing on this box: the oracle port of the same arithmetic
        kind = "port"
        from oracle import ssl_oracle as O
        w = random_init(cfg, 0)
        if cfg.family == "whisper":
            def one(x):
                hs = O.whisper_hidden_states(cfg, w, O.whisper_log_mel(w, x))
                f = O.select_features(hs, average=True)
                return O.masked_mean_pool(f, O.whisper_keep_frames(len(x), cfg.hidden_size))
        else:
            def one(x):
                return O.masked_mean_pool(O.select_features(O.w2v_hidden_states(cfg, w, x), average=True))

    def step():
        for x in waves:
            one(x)

    t_budget = time.time()
    warm_done = 0
    for _ in range(warmup):
        step()
        warm_done += 1
        if time.time() - t_budget > max_seconds / 3:
            break
    t0 = time.time()
    done = 0
    for _ in range(steps):
        step()
        done += 1
        if time.time() - t0 > max_seconds:
            break
    dt = time.time() - t0
    out = {"value": done * audio_per_step / dt, "unit": "audio-seconds/s", "cores": cores, "kind": kind,
           "sample": f"{done} step(s) x {n_utts} utterances x {wl['secs'] if wl['secs'] is not None else 'U[%g, %g]' % (wl['lo'], wl['hi'])} s, "
                     f"one utterance per forward (reference behaviour), fp32, torch.set_num_threads({cores})",
           "ms_per_step": 1e3 * dt / max(done, 1), "steps_done": done, "warmup_done": warm_done}
    if padded_too and batch_fn is not None:
        batch_fn(waves)
        t1 = time.time()
        reps = max(1, min(done, 3))
        for _ in range(reps):
            batch_fn(waves)
        out["padded_batch"] = {"value": reps * audio_per_step / (time.time() - t1), "unit": "audio-seconds/s",
                               "sample": f"{reps} forward(s) of one padded batch of the same {n_utts} utterances with attention_mask (BASELINE.md §3 mode ii)"}
    return out


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
class Timer:
    def __init__(self, torch, dist, world, dev):
        self.torch, self.dist, self.world, self.dev = torch, dist, world, dev
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def timed(self, fn, steps):
        torch = self.torch
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        for a, b in evs:
            self.flush.zero_()          # L2 flush, outside the timed events
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        wall = time.time() - t0
        self.barrier()
        ms_own = sum(a.elapsed_time(b) for a, b in evs)
        ms = ms_own
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall, ms_own


def cos_rel(a, b):
    import torch
    a, b = a.float().cpu().reshape(-1), b.float().cpu().reshape(-1)
    return (float(torch.nn.functional.cosine_similarity(a, b, dim=0)), float((a - b).abs().max() / b.abs().max()))


def golden_check(model, cfg):
    """The committed HuggingFace golden utterances (oracle/make_golden.py: HF fp32 forward on the same seed-0 weights)
    through the benched model: pooled mean-of-last-4 embedding, cosine and max|a - b| / max|b|."""
    import torch
    path = os.path.join(REPO, "tests", "golden", GOLDEN.get(cfg.name, ""))
    if not os.path.isfile(path):
        return None
    g = np.load(path)
    lens = [int(n) for n in g["lengths"]]
    waves = [synth_wave(int(g["wave_seed_base"]) + j, n) for j, n in enumerate(lens)]
    res = model.extract(waves, average=True, want_frames=False, want_pooled=True).pooled
    worst = (1.0, 0.0)
    for j in range(len(lens)):
        if f"meanlast4_pooled_{j}" in g.files:
            ref = torch.from_numpy(g[f"meanlast4_pooled_{j}"])
        else:   # Whisper fixture: pooled vectors of every hidden state; mean over layers and frames commute
            ref = torch.from_numpy(g[f"pooled_{j}"][-4:]).mean(0)
        c, r = cos_rel(res[j], ref)
        worst = (min(worst[0], c), max(worst[1], r))
    return {"utterances": len(lens), "samples": lens, "min_cosine": worst[0], "max_rel_err": worst[1], "fixture": "tests/golden/" + GOLDEN[cfg.name]}


def run_workload(name, args, ctx, model, steps, warmup, headline):
    """Times one workload on the already-built model. Returns (result dict for rank 0, parity ok)."""
    torch, dist, world, rank, dev, timer = ctx["torch"], ctx["dist"], ctx["world"], ctx["rank"], ctx["dev"], ctx["timer"]
    from interspeech_ser_b200 import scheduler
    wl = WORKLOADS[name]
    cfg = model.cfg
    eng = model.engine
    batch = args.batch if (args.batch and headline) else wl["batch"]
    strong = bool(wl.get("strong"))
    desc = wl["desc"]
    per_rank = {}

    if wl["secs"] is None:
        if strong:
            # ONE corpus for the whole job (seed 7); utterance i's waveform depends on i only, so every rank can
            # materialise exactly its own share and the gathered matrix is comparable across world sizes
            rng = np.random.default_rng(7)
            all_lens = [int(v) for v in rng.integers(int(wl["lo"] * 16000), int(wl["hi"] * 16000) + 1, size=batch)]
            plan_all, mine = scheduler.plan(cfg, all_lens, world, rank)
            plan = [plan_all[i] for i in mine]
            wave_of = lambda i: synth_wave(100000 + i, all_lens[i])  # noqa: E731
            secs_job = sum(all_lens) / 16000.0
        else:
            # corpus sweep: every rank owns its own synthetic corpus; the scheduler cuts it into packed batches
            rng = np.random.default_rng(7 + 1000 * rank)
            all_lens = [int(v) for v in rng.integers(int(wl["lo"] * 16000), int(wl["hi"] * 16000) + 1, size=batch)]
            plan = scheduler.make_batches(cfg, all_lens)
            cache = {}
            wave_of = lambda i: cache.setdefault(i, synth_wave(200000 + 1000003 * rank + i, all_lens[i]))  # noqa: E731
            secs_job = sum(all_lens) / 16000.0 * world
        hosts, devs, blens, bidx = [], [], [], []
        for bt in plan:
            ls = [all_lens[i] for i in bt.indices]
            hbuf = torch.from_numpy(np.concatenate([wave_of(i) for i in bt.indices])).pin_memory()
            hosts.append(hbuf); devs.append(hbuf.to(dev)); blens.append(ls); bidx.append(list(bt.indices))
        h2d_bytes = sum(h.numel() * 4 for h in hosts)
        n_local = sum(len(ls) for ls in blens)
        pooled_hosts = [torch.empty((len(ls), cfg.hidden_size), dtype=torch.float32).pin_memory() for ls in blens]
        last = {}

        def step_device():
            outs = []
            for wv, ls in zip(devs, blens):
                outs.append(model.extract_device(wv, ls, average=True, want_frames=False, want_pooled=True).pooled)
            last["pooled"] = outs
            return outs

        def step_e2e():
            for hb, ls, ph in zip(hosts, blens, pooled_hosts):
                out = model.extract_pinned(hb, ls, average=True, want_frames=False, want_pooled=True).pooled
                ph.copy_(out, non_blocking=True)

        def parity_rows():
            outs = last["pooled"]
            picks = [(0, 0), (len(outs) // 2, len(blens[len(outs) // 2]) // 2), (len(outs) - 1, len(blens[-1]) - 1)] if outs else []
            rows = []
            for bi, j in picks:
                i = bidx[bi][j]
                alone = model.extract_device(torch.from_numpy(wave_of(i)).to(dev), [all_lens[i]], average=True, use_graph=False).pooled[0]
                rows.append((f"utterance {i} ({all_lens[i] / 16000:.2f} s) of batch {bi}", outs[bi][j], alone))
            return rows
        desc += f"; {len(plan)} packed batches, {n_local} utterances, {sum(sum(ls) for ls in blens) / 16000.0:.0f} audio-s on this rank per step"
        d2h_bytes = n_local * cfg.hidden_size * 4
        per_rank = {"utterances": n_local, "batches": len(plan), "audio_s": sum(sum(ls) for ls in blens) / 16000.0}
    else:
        n = int(wl["secs"] * 16000)
        # every rank owns a different shard of the synthetic corpus (seed 7 = the scripts' default --seed)
        host = torch.from_numpy(synth_batch(7 + 1000 * rank, batch, n)).pin_memory()
        lens = [n] * batch
        host_flat = host.reshape(-1)
        wav_dev = host.to(dev).reshape(-1).contiguous()
        secs_job = batch * wl["secs"] * world
        h2d_bytes = host.numel() * 4
        d2h_bytes = batch * cfg.hidden_size * 4
        pooled_host = torch.empty((batch, cfg.hidden_size), dtype=torch.float32).pin_memory()
        last = {}

        def step_device():
            last["pooled"] = model.extract_device(wav_dev, lens, average=True, want_frames=False, want_pooled=True).pooled
            return last["pooled"]

        def step_e2e():
            # the public call for packed pinned input: upload through the model's two-slot ring on its copy stream (the
            # transfer of step i+1 runs under the encode of step i), encode, pooled rows back to pinned host memory
            out = model.extract_pinned(host_flat, lens, average=True, want_frames=False, want_pooled=True).pooled
            pooled_host.copy_(out, non_blocking=True)

        def parity_rows():
            rows = []
            for j in sorted({0, batch // 2, batch - 1}):
                alone = model.extract_device(wav_dev[j * n:(j + 1) * n].clone(), [n], average=True, use_graph=False).pooled[0]
                rows.append((f"row {j} of the benched batch", last["pooled"][j], alone))
            return rows

    n_graphs0 = len(eng._graphs)
    for _ in range(max(warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    # small fixed-shape batches: the public call replays a CUDA graph cached per length signature (engine.encode_w2v_graphed)
    graphed = len(eng._graphs) > n_graphs0

    sampler = ClockSampler(ctx["local_rank"])
    if rank == 0:
        sampler.start()
    ms_total, wall, ms_own = timer.timed(step_device, steps)
    clocks = sampler.stop() if rank == 0 else None

    # per-kernel-class device time for the roofline: separate, instrumented steps (an event pair around every launch)
    # run DIRECTLY behind the timed loop, so that they see the same power state / SM clock as the number they explain
    eng.set_profiling(True)
    prof_steps = 3 if headline else 1
    saved = os.environ.get("SERENC_NO_GRAPH")
    os.environ["SERENC_NO_GRAPH"] = "1"          # per-class events are recorded by the eager path only
    try:
        l0 = eng.launch_count()
        for _ in range(prof_steps):
            timer.flush.zero_()
            step_device()
        launches = (eng.launch_count() - l0) // prof_steps    # kernels per step (launched directly, or replayed from the cached graph)
        prof = eng.get_profile()
    finally:
        eng.set_profiling(False)
        if saved is None:
            os.environ.pop("SERENC_NO_GRAPH", None)
        else:
            os.environ["SERENC_NO_GRAPH"] = saved

    for _ in range(2):
        step_e2e()
    torch.cuda.synchronize()
    ms_e2e, _, _ = timer.timed(step_e2e, steps)

    # ---- parity inside the bench: benched rows vs the utterance alone, and the committed HF golden utterances ----
    step_device()
    rows = parity_rows()
    torch.cuda.synchronize()
    inv = {"rows": [], "bit_equal": True, "max_abs_diff": 0.0, "min_cosine": 1.0, "max_rel_err": 0.0}
    for what, got, alone in rows:
        c, r = cos_rel(got, alone)
        diff = float((got.float() - alone.float()).abs().max())
        inv["rows"].append(what)
        inv["bit_equal"] = inv["bit_equal"] and bool(torch.equal(got, alone))
        inv["max_abs_diff"] = max(inv["max_abs_diff"], diff)
        inv["min_cosine"], inv["max_rel_err"] = min(inv["min_cosine"], c), max(inv["max_rel_err"], r)
    gold = golden_check(model, cfg) if rank == 0 else None
    ok = inv["min_cosine"] >= COS_MIN and inv["max_rel_err"] <= REL_MAX and bool(torch.isfinite(rows[0][1]).all())
    if gold is not None:
        ok = ok and gold["min_cosine"] >= COS_MIN and gold["max_rel_err"] <= REL_MAX
    parity = {"ok": bool(ok), "tolerance": {"min_cosine": COS_MIN, "max_rel_err": REL_MAX},
              "benched_rows_vs_batch_of_one": inv, "hf_golden_through_benched_model": gold}

    # ---- strong scaling: gather the pooled matrix (the path's only exchange), un-permute, checksum ----
    extra = {}
    if strong:
        mine = {}
        for idxs, out in zip(bidx, last["pooled"]):
            oc = out.cpu()
            for j, i in enumerate(idxs):
                mine[i] = oc[j]
        gathered = [None] * world if rank == 0 else None
        if world > 1:
            dist.gather_object(mine, gathered, dst=0)
        else:
            gathered = [mine]
        busy = torch.tensor([ms_own / steps], dtype=torch.float64, device=dev)
        busy_all = [torch.zeros_like(busy) for _ in range(world)]
        if world > 1:
            dist.all_gather(busy_all, busy)
        else:
            busy_all = [busy]
        if rank == 0:
            mat = torch.stack(scheduler.merge_rank_results(gathered, len(all_lens)))
            extra = {"gathered_shape": list(mat.shape), "gathered_sha256": hashlib.sha256(mat.numpy().tobytes()).hexdigest(),
                     "per_rank_busy_ms_per_step": [float(b.item()) for b in busy_all],
                     "note": "the sha256 of the gathered, un-permuted [utterances, d] matrix is the same for every world size (bitwise equality with N = 1)"}

    if rank != 0:
        return None, ok
    value = secs_job * steps / (ms_total / 1e3)
    e2e_value = secs_job * steps / (ms_e2e / 1e3)
    peaks = measured_peaks()
    gemm_ms = sum(prof[k]["ms"] for k in eng.GEMM_CLASSES)
    gemm_fl = sum(prof[k]["flops"] for k in eng.GEMM_CLASSES)
    gemm_n = sum(prof[k]["launches"] for k in eng.GEMM_CLASSES)
    total_prof_ms = sum(v["ms"] for v in prof.values())
    achieved = gemm_fl / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    peak = peaks["bf16_tflops_sustained"]
    breakdown = {k: {"ms_per_step": v["ms"] / prof_steps, "share": v["ms"] / total_prof_ms if total_prof_ms else 0.0,
                     "launches_per_step": v["launches"] // prof_steps,
                     "tflops": (v["flops"] / (v["ms"] / 1e3) / 1e12) if v["ms"] > 0 and v["flops"] > 0 else None,
                     "gbs": (v["bytes"] / (v["ms"] / 1e3) / 1e9) if v["ms"] > 0 and v["bytes"] > 0 else None}
                 for k, v in prof.items() if v["launches"]}
    total_flops = sum(v["flops"] for v in prof.values()) / prof_steps
    cfgd = workload_config(name, world, batch)
    cfgd["workload"] = desc
    cfgd["cuda_graph"] = bool(graphed)
    res = {
        "metric": metric_name(name),
        "value": value, "unit": "audio-seconds/s", "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
        "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": cfgd,
        "e2e": {"value": e2e_value, "unit": "audio-seconds/s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                "ms_per_step": ms_e2e / steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "gemm_bf16_tcgen05_2cta_kernel / gemm_bf16_tcgen05_kernel (all linear + implicit-GEMM conv launches of a step)", "bound": "tensor",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                     "frac_of_burst_peak": achieved / peaks["bf16_tflops"] if peaks["bf16_tflops"] else None,
                     "traffic": ncu_traffic(name), "traffic_source": "static: committed ncu capture (profiles/ncu_traffic.json), not measured by this run" if ncu_traffic(name) else None,
                     "peak_source": f"{peaks['source']} (bf16_tflops_sustained; burst {peaks['bf16_tflops']})",
                     "launches_per_step": gemm_n // prof_steps, "share_of_step": gemm_ms / total_prof_ms if total_prof_ms else None,
                     "algorithmic_gflop_per_step": gemm_fl / prof_steps / 1e9},
        "model_tflops": total_flops / (ms_total / steps / 1e3) / 1e12,   # per GPU
        "kernel_breakdown": breakdown,
        "parity_check": parity,
        "wall_s_timed_region": wall,
    }
    if per_rank:
        res["rank0_share"] = per_rank
    if extra:
        res["strong_scaling"] = extra
    return res, ok


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")

    from interspeech_ser_b200 import configs
    from interspeech_ser_b200.modeling import SpeechEncoderModel, WhisperModel
    from interspeech_ser_b200.weights import random_init

    extras = []
    if args.workloads != "none" and args.workload == "wavlm-large" and not args.graph:
        extras = DEFAULT_EXTRAS if args.workloads == "all" else [w for w in args.workloads.split(",") if w]
        for w in extras:
            if w not in WORKLOADS:
                raise SystemExit(f"unknown workload {w!r}")
    names = [args.workload] + [w for w in extras if w != args.workload]

    # Random-init weights are generated on the host (up to 2.2 G parameters: ~50 s of PCG64 for XLS-R-2b); the
    # generator releases the GIL, so the models of the later workloads are generated while the earlier ones run.
    model_order = []
    for n in names:
        if WORKLOADS[n]["model"] not in model_order:
            model_order.append(WORKLOADS[n]["model"])
    # One rank: all at once. Several ranks share the host's memory (XLS-R-2b is 8.6 GB of fp32 per rank): one model ahead.
    pool = ThreadPoolExecutor(max_workers=max(1, len(model_order)))
    futures = {}

    def prefetch(upto):
        for m in model_order[:upto]:
            if m not in futures and m not in consumed:
                futures[m] = pool.submit(random_init, configs.get_config(m), 0)
    consumed = set()
    prefetch(len(model_order) if world == 1 else 1)

    ctx = {"torch": torch, "dist": dist, "world": world, "rank": rank, "local_rank": local_rank, "dev": dev,
           "timer": Timer(torch, dist, world, dev), "graph_launches": {}}
    t_start = time.time()
    results, all_ok = {}, True
    for mi, m in enumerate(model_order):
        cfg = configs.get_config(m)
        prefetch(mi + 2)
        weights = futures.pop(m).result()
        consumed.add(m)
        model = (WhisperModel if cfg.family == "whisper" else SpeechEncoderModel)(cfg, weights, local_rank)
        del weights
        for n in names:
            if WORKLOADS[n]["model"] != m:
                continue
            headline = n == args.workload
            steps = args.steps if headline else max(1, min(args.steps, args.extra_steps))
            res, ok = run_workload(n, args, ctx, model, steps, args.warmup, headline)
            all_ok = all_ok and ok
            if res is not None:
                res["wall_s_since_start"] = time.time() - t_start
                results[n] = res
        model.engine.close()
        del model
        torch.cuda.empty_cache()
    pool.shutdown()
    if world > 1:
        flag = torch.tensor([0 if all_ok else 1], device=dev)
        dist.all_reduce(flag)
        all_ok = int(flag.item()) == 0

    if rank == 0:
        line = results[args.workload]
        if extras:
            keep = ("value", "unit", "ms_per_step", "steps", "warmup", "scaling", "e2e", "gpu_launches", "roofline", "kernel_breakdown",
                    "parity_check", "config", "model_tflops", "clocks")
            line["workloads"] = {n: {k: results[n][k] for k in keep if k in results[n]} for n in names if n != args.workload and n in results}
        line["parity_ok"] = bool(all_ok)
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb = cpu_reference_run(args.workload, steps=3, warmup=1, n_utts=8 if "whisper" not in args.workload else 2, max_seconds=40.0)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": "audio-seconds/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not all_ok:
        sys.stderr.write("bench.py: parity_check FAILED (see parity_check in the JSON line)\n")
        raise SystemExit(3)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    n_utts = 8 if "whisper" not in args.workload else 2
    res = cpu_reference_run(args.workload, steps=args.steps, warmup=args.warmup, n_utts=n_utts, max_seconds=150.0, padded_too=True)
    cfgd = workload_config(args.workload, world)
    cfgd["reference_sample"] = ("reference's own CPU implementation (HF transformers forward, one utterance per forward) on the host cores; "
                                f"each step is a bounded sample of the workload: {n_utts} of its utterances")
    line = {
        "impl": "reference",
        "metric": metric_name(args.workload),
        "value": res["value"], "unit": "audio-seconds/s", "n_gpus": world, "steps": res["steps_done"], "warmup": res["warmup_done"],
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfgd,
        "cpu_baseline": {"value": res["value"], "unit": "audio-seconds/s", "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": "audio-seconds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if res["steps_done"] < args.steps or res["warmup_done"] < args.warmup:
        line["note"] = (f"stopped after {res['warmup_done']} of {args.warmup} warm-up and {res['steps_done']} of {args.steps} timed steps: "
                        "the CPU arm is bounded to ~150 s of timed work")
    if "padded_batch" in res:
        line["padded_batch"] = res["padded_batch"]
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="wavlm-large", choices=sorted(WORKLOADS), help="the headline workload (top-level keys of the JSON line)")
    ap.add_argument("--workloads", default="all",
                    help="other workloads timed in the same run and reported under `workloads`: all (the BASELINE configs) | none | comma list; "
                         "only with the default headline")
    ap.add_argument("--extra-steps", type=int, default=5, help="timed steps of each additional workload (at most --steps)")
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch of the headline workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", action="store_true", help="(kept for compatibility: the public call now replays small batches from a cached CUDA graph by itself)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
            # `python bench.py --gpus N` typed by hand: become the launch the driver uses (one rank per GPU, rendezvous on 127.0.0.1)
            import socket
            with socket.socket() as sk:
                sk.bind(("127.0.0.1", 0))
                port = sk.getsockname()[1]
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
            os.execvp(cmd[0], cmd)
        run_ours(args)


if __name__ == "__main__":
    main()
