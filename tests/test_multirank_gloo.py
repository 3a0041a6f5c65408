"""World-size-2 test of the N>1 host logic on CPU (gloo): deterministic sharding of length-bucketed batches and the
final host gather. The data path itself has no collective (replicas run identical kernels on disjoint utterances)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from interspeech_ser_b200 import configs, scheduler


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fake_embed(length: int, d: int = 8) -> torch.Tensor:
    # stand-in for the per-utterance pooled embedding: depends only on the utterance itself (batching invariance)
    g = torch.Generator().manual_seed(int(length))
    return torch.randn(d, generator=g)


def _worker(rank, world, port, lens, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = configs.get_config("microsoft/wavlm-large")
    batches, mine = scheduler.plan(cfg, lens, world, rank, frame_budget=4096)
    rows = {}
    for bi in mine:
        for i in batches[bi].indices:
            rows[i] = _fake_embed(lens[i])
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(rows, gathered, dst=0)
    if rank == 0:
        merged = scheduler.merge_rank_results(gathered, len(lens))
        torch.save(torch.stack(merged), os.path.join(out_dir, "pooled.pt"))
        torch.save([sorted(g) for g in gathered], os.path.join(out_dir, "owners.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path):
    rng = np.random.default_rng(5)
    lens = [int(v) for v in rng.integers(32000, 192000, size=64)]
    port = _free_port()
    mp.spawn(_worker, args=(2, port, lens, str(tmp_path)), nprocs=2, join=True)
    pooled = torch.load(tmp_path / "pooled.pt")
    owners = torch.load(tmp_path / "owners.pt")
    ref = torch.stack([_fake_embed(n) for n in lens])
    assert torch.equal(pooled, ref)                     # same matrix as a single rank would produce, in corpus order
    assert sorted(owners[0] + owners[1]) == list(range(64)) and owners[0] and owners[1]
    assert not set(owners[0]) & set(owners[1])


def _cli_shard_worker(rank, world, port, sizes, lens, out_dir):
    """The CLI's N>1 path: ranks shard FILES by size on disk (no decode of other ranks' files), run their windows
    through make_batches locally, and rank 0 gathers {name: pooled row}."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = configs.get_config("microsoft/wavlm-large")
    mine = scheduler.shard_by_cost(sizes, world)[rank]
    rows = {}
    for w0 in range(0, len(mine), 5):                     # decode windows of 5 files
        window = mine[w0:w0 + 5]
        for bt in scheduler.make_batches(cfg, [lens[i] for i in window], frame_budget=2048):
            for j in bt.indices:
                rows[window[j]] = _fake_embed(lens[window[j]])
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(rows, gathered, dst=0)
    if rank == 0:
        torch.save(torch.stack(scheduler.merge_rank_results(gathered, len(lens))), os.path.join(out_dir, "pooled_cli.pt"))
        torch.save([sum(sizes[i] for i in g) for g in gathered], os.path.join(out_dir, "loads.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_cli_file_sharding(tmp_path):
    rng = np.random.default_rng(9)
    lens = [int(v) for v in rng.integers(32000, 320000, size=41)]
    sizes = [44.0 + 2.0 * n for n in lens]                # PCM16 WAV: header + 2 bytes per sample
    port = _free_port()
    mp.spawn(_cli_shard_worker, args=(2, port, sizes, lens, str(tmp_path)), nprocs=2, join=True)
    pooled = torch.load(tmp_path / "pooled_cli.pt")
    assert torch.equal(pooled, torch.stack([_fake_embed(n) for n in lens]))
    loads = torch.load(tmp_path / "loads.pt")
    assert abs(loads[0] - loads[1]) <= max(sizes)
