#!/usr/bin/env python
"""Benchmark of the embedding-extraction hot path: audio-seconds encoded per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload wavlm-large|whisper-large-v3|hubert-xlarge|xls-r-2b]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...     # the reference's own CPU implementation (HF transformers) on host cores

One "step" = one pass of the hot path over one batch of synthetic utterances per GPU:
waveforms -> (normalise ->) encoder -> mean of the last four hidden states (preprocess_speech.py:56-63,
`--use_average y`) -> masked-mean pooled embedding per utterance.

  value   device-resident inputs, device-timed (CUDA events), L2 flushed between steps (untimed)
  e2e     pinned host waveforms -> H2D -> encode -> pooled [B, d] -> D2H, through the public Python API
Weak scaling: every rank encodes its own shard of the batch; no data-path collective, one final gather.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WAVE_STD = 0.0886  # MSP-Podcast corpus std (reference: benchmark/model/cat_ser/7/train_norm_stat.pkl)

WORKLOADS = {
    # name: (config, utterance seconds, per-GPU batch, description)
    # 142 x 199 frames = 28 258 rows = 111 row-tiles of 256: every transformer GEMM is then a whole number of
    # waves over the 74 CTA pairs (111 x {4, 12, 16} n-tiles = {6, 18, 24} x 74) — the scheduler's frame budget.
    "wavlm-large": ("microsoft/wavlm-large", 4.0, 142,
                    "WavLM-large (random-init) embedding extraction, 4 s synthetic 16 kHz utterances "
                    "(BASELINE configs[0] utterance shape), batch 142 per GPU"),
    "wavlm-large-c1": ("microsoft/wavlm-large", 4.0, 8,
                       "WavLM-large (random-init) embedding extraction, batch 8 x 4 s synthetic 16 kHz utterances (BASELINE configs[0])"),
    "whisper-large-v3": ("openai/whisper-large-v3", 30.0, 32,
                         "Whisper-large-v3 encoder: 128-bin log-mel frontend + encoder over 30 s synthetic audio, batch 32 per GPU (BASELINE configs[1])"),
    "hubert-xlarge": ("facebook/hubert-xlarge-ls960-ft", 8.0, 64, "HuBERT-xlarge-ls960 embedding extraction, 8 s utterances, batch 64 per GPU"),
    "xls-r-2b": ("facebook/wav2vec2-xls-r-2b", 8.0, 64, "wav2vec2-xls-r-2b embedding extraction, batch 64 x 8 s per GPU (BASELINE configs[3])"),
    # corpus sweep in the style of BASELINE configs[4]: a step = the whole per-GPU corpus, length-sorted into packed
    # batches by the scheduler (frame budget 28 416), one encode call per batch
    "wavlm-large-sweep": ("microsoft/wavlm-large", None, 512,
                          "WavLM-large (random-init) corpus sweep: 512 synthetic utterances per GPU, lengths U[2 s, 20 s], "
                          "length-sorted packed batches from the scheduler (BASELINE configs[4] style)"),
}


def synth_batch(seed: int, batch: int, n_samples: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((batch, n_samples), dtype=np.float32) * np.float32(WAVE_STD)).astype(np.float32)


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


# --------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: HF transformers fp32 on the host cores (what the reference scripts call)
# --------------------------------------------------------------------------------------------------
def ncu_traffic(workload: str):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/ncu_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum averaged over the four GEMMs of one WavLM-large encoder layer).
    A static, profiler-side number: null for workloads that were not captured."""
    if workload != "wavlm-large":
        return None
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        return None


def cpu_reference_run(workload: str, steps: int, warmup: int, n_utts: int, max_seconds: float = 150.0):
    """One utterance per forward, as preprocess_speech.py:45-73 / preprocess_whisper.py:45-82 do."""
    import torch

    cfg_name, secs, _, _ = WORKLOADS[workload]
    from interspeech_ser_b200 import configs
    from interspeech_ser_b200.weights import random_init

    cfg = configs.get_config(cfg_name)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if secs is None:
        secs = 11.0   # corpus sweep: mean utterance length of U[2 s, 20 s]
    n = int(secs * 16000)
    waves = synth_batch(7, n_utts, n)
    kind = "reference"
    try:
        import transformers  # noqa: F401
        from oracle.make_golden import hf_model
        import transformers as tr
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
        else:
            fe = tr.Wav2Vec2FeatureExtractor(feature_size=1, sampling_rate=16000, padding_value=0.0, do_normalize=True, return_attention_mask=True)

            def one(x):
                inputs = fe(x, sampling_rate=16000, return_tensors="pt", padding=True)
                with torch.no_grad():
                    hs = model(**inputs, output_hidden_states=True).hidden_states
                return torch.mean(torch.stack(hs[-4:]), dim=0).squeeze(0).mean(0)
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
        for b in range(n_utts):
            one(waves[b])

    t_budget = time.time()
    for _ in range(warmup):
        step()
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
    audio_s = done * n_utts * secs
    return {"value": audio_s / dt, "unit": "audio-seconds/s", "cores": cores, "kind": kind,
            "sample": f"{done} step(s) x {n_utts} utterances x {secs:g} s, one utterance per forward (reference behaviour), fp32, "
                      f"torch.set_num_threads({cores})", "ms_per_step": 1e3 * dt / max(done, 1), "steps_done": done}


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
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

    cfg_name, secs, batch, desc = WORKLOADS[args.workload]
    if args.batch:
        batch = args.batch
    cfg = configs.get_config(cfg_name)
    weights = random_init(cfg, 0)
    model = (WhisperModel if cfg.family == "whisper" else SpeechEncoderModel)(cfg, weights, local_rank)
    del weights
    eng = model.engine
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    if secs is None:
        # corpus sweep: every rank owns its own synthetic corpus; the scheduler cuts it into packed batches
        from interspeech_ser_b200 import scheduler
        rng = np.random.default_rng(7 + 1000 * rank)
        all_lens = [int(v) for v in rng.integers(2 * 16000, 20 * 16000 + 1, size=batch)]
        plan = scheduler.make_batches(cfg, all_lens)
        hosts, devs, blens = [], [], []
        for bi, bt in enumerate(plan):
            ls = [all_lens[i] for i in bt.indices]
            hbuf = torch.from_numpy((rng.standard_normal(sum(ls), dtype=np.float32) * np.float32(WAVE_STD))).pin_memory()
            hosts.append(hbuf); devs.append(hbuf.to(dev)); blens.append(ls)
        secs_total = sum(all_lens) / 16000.0
        h2d_bytes = sum(h.numel() * 4 for h in hosts)
        pooled_hosts = [torch.empty((len(ls), cfg.hidden_size), dtype=torch.float32).pin_memory() for ls in blens]

        def step_device():
            out = None
            for wv, ls in zip(devs, blens):
                out = model.extract_device(wv, ls, average=True, want_frames=False, want_pooled=True).pooled
            return out

        def step_e2e():
            for hb, ls, ph in zip(hosts, blens, pooled_hosts):
                out = model.extract_pinned(hb, ls, average=True, want_frames=False, want_pooled=True).pooled
                ph.copy_(out, non_blocking=True)

        pooled_host = pooled_hosts[-1]
        host = None
        desc += f"; {len(plan)} batches, {sum(len(b) for b in blens)} utterances, {secs_total:.0f} audio-s per GPU and step"
    else:
        n = int(secs * 16000)
        # every rank owns a different shard of the synthetic corpus (seed 7 = the scripts' default --seed)
        host = torch.from_numpy(synth_batch(7 + 1000 * rank, batch, n)).pin_memory()
        lens = [n] * batch
        host_flat = host.reshape(-1)
        wav_dev = host.to(dev).reshape(-1).contiguous()
        secs_total = batch * secs
        h2d_bytes = host.numel() * 4

        def step_device():
            return model.extract_device(wav_dev, lens, average=True, want_frames=False, want_pooled=True).pooled

        pooled_host = torch.empty((batch, cfg.hidden_size), dtype=torch.float32).pin_memory()

        def step_e2e():
            # the public call for packed pinned input: upload through the model's two-slot ring on its copy stream (the
            # transfer of step i+1 runs under the encode of step i), encode, pooled rows back to pinned host memory
            out = model.extract_pinned(host_flat, lens, average=True, want_frames=False, want_pooled=True).pooled
            pooled_host.copy_(out, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()

    def timed(fn, steps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        for a, b in evs:
            flush.zero_()          # L2 flush, outside the timed events
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        wall = time.time() - t0
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall

    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    graph_launches = None
    if args.graph:
        if secs is None:
            raise SystemExit("--graph needs a fixed-shape workload")
        # the whole encode call (span upload, ~200 kernels, pooling) captured once through the C ABI, replayed per step
        graph = torch.cuda.CUDAGraph()
        lg = eng.launch_count()
        with torch.cuda.graph(graph):
            graph_out = step_device()
        graph_launches = eng.launch_count() - lg
        eager_step = step_device

        def step_device():
            graph.replay()
            return graph_out
        step_device()
        torch.cuda.synchronize()
        if not torch.equal(graph_out, eager_step()):
            raise SystemExit("CUDA-graph replay differs from the eager call")
    l0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, wall = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = graph_launches if graph_launches is not None else (eng.launch_count() - l0) // args.steps

    for _ in range(2):
        step_e2e()
    torch.cuda.synchronize()
    ms_e2e, _ = timed(step_e2e, args.steps)

    # final host gather of the pooled embeddings (the path's only exchange)
    if world > 1 and secs is not None:
        gathered = [torch.empty_like(pooled_host, device=dev) for _ in range(world)] if rank == 0 else None
        dist.gather(pooled_host.to(dev), gathered, dst=0)

    # per-kernel-class device time for the roofline (separate, instrumented steps)
    eng.set_profiling(True)
    prof_steps = 2
    if args.graph:
        step_device = eager_step   # per-class events are recorded by the eager path only
    for _ in range(prof_steps):
        flush.zero_()
        step_device()
    prof = eng.get_profile()
    eng.set_profiling(False)

    audio_per_step = secs_total * world
    value = audio_per_step * args.steps / (ms_total / 1e3)
    e2e_value = audio_per_step * args.steps / (ms_e2e / 1e3)

    if rank == 0:
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
        line = {
            "metric": "audio-seconds/sec encoded" + (" (WavLM-large)" if "wavlm" in args.workload else f" ({cfg.name})"),
            "value": value, "unit": "audio-seconds/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": desc, "model": cfg.name, "per_gpu_batch": batch, "utterance_seconds": secs,
                       "global_batch": batch * world, "output": "mean of last 4 hidden states -> masked-mean pooled [B, d] fp32",
                       "weights": "random init (seed 0)", "l2": "256 MiB buffer zeroed between timed steps (untimed)",
                       "parallelism": f"utterance-sharded replicas x{world}, no data-path collective",
                       "cuda_graph": bool(args.graph)},
            "e2e": {"value": e2e_value, "unit": "audio-seconds/s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(batch * cfg.hidden_size * 4),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "gemm_bf16_tcgen05_kernel (all linear + implicit-GEMM conv launches of a step)", "bound": "tensor",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                         "traffic": ncu_traffic(args.workload), "peak_source": f"{peaks['source']} (bf16_tflops_sustained; burst {peaks['bf16_tflops']})",
                         "launches_per_step": gemm_n // prof_steps, "share_of_step": gemm_ms / total_prof_ms if total_prof_ms else None,
                         "algorithmic_gflop_per_step": gemm_fl / prof_steps / 1e9},
            "model_tflops": total_flops / (ms_total / args.steps / 1e3) / 1e12,
            "kernel_breakdown": breakdown,
            "wall_s_timed_region": wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb = cpu_reference_run(args.workload, steps=1, warmup=0, n_utts=8 if cfg.family != "whisper" else 2, max_seconds=40.0)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": "audio-seconds/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cfg_name, secs, batch, desc = WORKLOADS[args.workload]
    n_utts = 8 if "whisper" not in args.workload else 2
    res = cpu_reference_run(args.workload, steps=args.steps, warmup=min(args.warmup, 1), n_utts=n_utts, max_seconds=150.0)
    line = {
        "impl": "reference",
        "metric": "audio-seconds/sec encoded" + (" (WavLM-large)" if "wavlm" in args.workload else f" ({cfg_name})"),
        "value": res["value"], "unit": "audio-seconds/s", "n_gpus": world, "steps": res["steps_done"], "warmup": min(args.warmup, 1),
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "model": cfg_name, "per_gpu_batch": n_utts, "utterance_seconds": secs,
                   "note": "reference's own CPU implementation (HF transformers forward, one utterance per forward) on the host cores; "
                           "each step is a bounded sample of the workload"},
        "cpu_baseline": {"value": res["value"], "unit": "audio-seconds/s", "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": "audio-seconds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="wavlm-large", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", action="store_true",
                    help="replay the device-resident step from a CUDA graph (fixed-shape workloads; pays off at small batches)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
