"""Length bucketing and utterance sharding (host logic, no GPU needed).

The reference processes one file per forward from 4 Python threads (preprocess_speech.py:120-122). Here
utterances are sorted by length, cut into batches under a frame budget (packed layout => no padding inside a
batch), and whole batches are dealt to the GPUs of one box by longest-processing-time-first on the analytic FLOP
model of SURVEY.md §8d. The path is embarrassingly data-parallel: no collective besides the final gather.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

from .configs import ARCH_WHISPER, EncoderConfig, w2v_num_frames


def utterance_flops(cfg: EncoderConfig, n_samples: int) -> float:
    """Algorithmic FLOPs of one utterance:
    FE + 2 T 512 d + 2 T d (d/g) k_pos + L 2 T (4 d^2 + 2 d ffn) + L 4 T^2 d   (Whisper: fixed 30 s window)."""
    d, L, ffn = cfg.hidden_size, cfg.num_hidden_layers, cfg.intermediate_size
    if cfg.arch == ARCH_WHISPER:
        T = 1500
        conv = 2.0 * 3000 * cfg.num_mel_bins * d * 3 + 2.0 * 1500 * d * d * 3
        return conv + L * 2.0 * T * (4 * d * d + 2 * d * ffn) + L * 4.0 * T * T * d
    n = int(n_samples)
    fe = 0.0
    cin = 1
    for k, s, c in zip(cfg.conv_kernel, cfg.conv_stride, cfg.conv_dim):
        n = (n - k) // s + 1 if n >= k else 0
        fe += 2.0 * cin * c * k * n
        cin = c
    T = n
    proj = 2.0 * T * cfg.conv_dim[-1] * d
    pos = 2.0 * T * d * (d // cfg.num_conv_pos_embedding_groups) * cfg.num_conv_pos_embeddings
    return fe + proj + pos + L * 2.0 * T * (4 * d * d + 2 * d * ffn) + L * 4.0 * T * T * d


@dataclass
class Batch:
    indices: List[int]      # positions in the caller's utterance list
    frames: int             # sum of frames in the batch
    flops: float


def make_batches(cfg: EncoderConfig, lengths: Sequence[int], frame_budget: int = 28416, max_batch: int = 1024,
                 max_spread: float = 0.0) -> List[Batch]:
    """Sort by length and cut into batches of at most `frame_budget` frames / `max_batch` utterances.
    The default budget, 28 416 = 111 x 256 frames, makes the row-tile count of the 256 x 256 CTA-pair GEMM a
    multiple of 37, so that every transformer GEMM (4 / 12 / 16 column tiles) is a whole number of waves over the
    74 CTA pairs of a B200.
    max_spread > 0 additionally closes a batch when the longest/shortest frame ratio would exceed 1 + max_spread
    (length bucketing; with the packed layout padding costs nothing, so spread only matters for attention tiles).
    Utterances too short to yield a frame are rejected up front (HF would raise inside the conv stack)."""
    def frames_of(n):
        return 1500 if cfg.arch == ARCH_WHISPER else w2v_num_frames(n, cfg)

    order = sorted(range(len(lengths)), key=lambda i: (lengths[i], i))
    batches: List[Batch] = []
    cur: List[int] = []
    cur_frames, cur_flops, cur_min = 0, 0.0, 0
    for i in order:
        t = frames_of(lengths[i])
        if t < 1:
            raise ValueError(f"utterance {i}: {lengths[i]} samples yields no frame (needs >= 400 samples)")
        close = bool(cur) and (cur_frames + t > frame_budget or len(cur) >= max_batch or
                               (max_spread > 0 and t > cur_min * (1.0 + max_spread)))
        if close:
            batches.append(Batch(cur, cur_frames, cur_flops))
            cur, cur_frames, cur_flops = [], 0, 0.0
        if not cur:
            cur_min = t
        cur.append(i)
        cur_frames += t
        cur_flops += utterance_flops(cfg, lengths[i])
    if cur:
        batches.append(Batch(cur, cur_frames, cur_flops))
    return batches


def shard_batches(batches: Sequence[Batch], world_size: int) -> List[List[int]]:
    """Greedy LPT: heaviest batch first onto the least-loaded rank. Deterministic (ties -> lower rank)."""
    loads = [0.0] * world_size
    assign: List[List[int]] = [[] for _ in range(world_size)]
    for bi in sorted(range(len(batches)), key=lambda j: (-batches[j].flops, j)):
        r = min(range(world_size), key=lambda k: (loads[k], k))
        assign[r].append(bi)
        loads[r] += batches[bi].flops
    for a in assign:
        a.sort()
    return assign


def shard_by_cost(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Greedy LPT over single items (e.g. files, cost = size on disk as a proxy for audio length): heaviest first onto
    the least-loaded rank. Each rank's list comes back sorted by cost (then index), so that consecutive windows of it
    hold utterances of similar length. Deterministic; every index appears on exactly one rank."""
    loads = [0.0] * world_size
    assign: List[List[int]] = [[] for _ in range(world_size)]
    for i in sorted(range(len(costs)), key=lambda j: (-costs[j], j)):
        r = min(range(world_size), key=lambda k: (loads[k], k))
        assign[r].append(i)
        loads[r] += costs[i]
    for a in assign:
        a.sort(key=lambda j: (costs[j], j))
    return assign


def plan(cfg: EncoderConfig, lengths: Sequence[int], world_size: int, rank: int, **kw) -> Tuple[List[Batch], List[int]]:
    """Batches for the whole corpus and the indices of the batches this rank runs."""
    batches = make_batches(cfg, lengths, **kw)
    return batches, shard_batches(batches, world_size)[rank]


def merge_rank_results(per_rank: Sequence[Dict[int, object]], n_total: int) -> List[object]:
    """Final host gather: every rank returns {utterance index: result}; un-permute into corpus order."""
    out: List[object] = [None] * n_total
    for res in per_rank:
        for i, v in res.items():
            if out[i] is not None:
                raise ValueError(f"utterance {i} produced by two ranks")
            out[i] = v
    missing = [i for i, v in enumerate(out) if v is None]
    if missing:
        raise ValueError(f"{len(missing)} utterances missing from the gather, first {missing[:5]}")
    return out
