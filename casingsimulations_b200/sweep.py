"""Multi-GPU partitioning of independent solves.

Frequencies (and DC / TDEM sources, model variants) are independent iterations of
SimPEG's loop (``for freq in survey.freqs``; casingSimulations creates one source
per frequency, sources.py:119-123), so they shard across ranks with no collective
on the data path; one all-gather of the solutions at the end is optional.  One
process per GPU, torch.distributed for the plumbing (nccl on GPUs, gloo in tests).
"""
import numpy as np


def shard_indices(n, rank, world):
    """Round-robin: low frequencies are the expensive ones for an iterative solver and
    neighbours in a log-spaced sweep cost the same, so interleaving balances the load."""
    return np.arange(rank, n, world)


def shard_frequencies(freqs, rank, world):
    freqs = np.asarray(freqs, dtype=float)
    return freqs[shard_indices(len(freqs), rank, world)]


def gather_solutions(local, n_total, rank, world, group=None):
    """All-gather per-rank solution blocks (n_local, n) into the (n_total, n) array in the
    original frequency order.  ``local`` is a torch tensor on the backend's device."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    counts = [len(shard_indices(n_total, r, world)) for r in range(world)]
    nmax, n = max(counts), local.shape[1]
    pad = torch.zeros((nmax, n), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    if local.is_complex():
        buf = [torch.zeros((nmax, n, 2), dtype=torch.float64, device=local.device) for _ in range(world)]
        dist.all_gather(buf, torch.view_as_real(pad).contiguous(), group=group)
        buf = [torch.view_as_complex(b) for b in buf]
    else:
        buf = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(buf, pad, group=group)
    out = torch.zeros((n_total, n), dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = torch.as_tensor(shard_indices(n_total, r, world), device=local.device)
        out[idx] = buf[r][: counts[r]]
    return out
