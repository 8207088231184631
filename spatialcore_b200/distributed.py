"""Multi-GPU sharding of the permutation null (one process per GPU, ``torch.distributed``).

Permutations are independent given (W, Z, lag): every rank holds the graph and the standardised
matrices, runs a contiguous block of the P permutations and the per-gene null summaries
(count_ge, count_abs_ge, Σsim, Σsim²; a [4, G] FP64 tensor, 32 KB at G = 1000) are summed with ONE
all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests).  Philox permutations are addressed by
their global index, so the result does not depend on the world size.  There is no other exchange
on this path, hence no fused compute+collective kernel.  Gene blocks are the other natural partition
(each rank owns G/W genes end to end, one all-gather of per-gene vectors); ``gene_groups_for`` picks
the hybrid of the two that keeps matrix rows wide.
"""

from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def block_slice(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of ``range(total)``: sizes differ by at most one."""
    base, rem = divmod(total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def my_slice(total: int, group=None) -> Tuple[int, int]:
    rank, ws = world(group)
    return block_slice(total, rank, ws)


def gene_groups_for(n_genes: int, world_size: int, min_genes: int = 500) -> int:
    """Hybrid partition: the largest divisor d of the world size that keeps >= ``min_genes`` genes per
    block (rows of the N x G matrices stay wide enough for the gather kernel to run at HBM speed);
    the remaining factor world/d shards the permutations."""
    best = 1
    for d in range(1, world_size + 1):
        if world_size % d == 0 and n_genes // d >= min_genes:
            best = d
    return best


def combine_moments(counts, means, stds):
    """Pool per-block (count, mean, population std) of each gene into the moments of the whole
    matrix (Chan et al. pairwise update, FP64 numpy).  Inputs are [B], [B, G], [B, G].
    Differences are taken against block 0's mean so that a column that is constant everywhere pools
    to variance exactly 0 (it must be flagged zero-variance, not divided by 1e-17)."""
    import numpy as np

    counts = np.asarray(counts, dtype=np.float64).reshape(-1, 1)
    means = np.asarray(means, dtype=np.float64)
    stds = np.asarray(stds, dtype=np.float64)
    total = counts.sum()
    delta0 = means - means[0:1]
    mean = means[0] + (counts * delta0).sum(0) / total
    m2 = (counts * stds * stds).sum(0) + (counts * (means - mean[None, :]) ** 2).sum(0)
    const = np.all(stds == 0, axis=0) & np.all(delta0 == 0, axis=0)
    var = np.where(const, 0.0, m2 / total)
    std = np.sqrt(var)
    return mean, std, (std == 0)


def row_block(n: int, rank: int, world_size: int) -> Tuple[int, int, int]:
    """Equal-sized row blocks for ``all_gather_into_tensor``: returns ``(rows_per_rank, lo, hi)``;
    the last blocks may be short or empty."""
    per = (n + world_size - 1) // world_size
    lo = min(n, rank * per)
    return per, lo, min(n, lo + per)


def bind_host_to_gpu(device_index: int) -> bool:
    """Restrict this process to the CPUs local to the GPU's PCIe root complex (its NUMA node), so that
    host buffers pinned afterwards are NUMA-local and eight ranks uploading at once do not funnel through
    one socket's memory and the inter-socket link.  Returns False (and changes nothing) when the
    topology cannot be read."""
    import os

    try:
        props = torch.cuda.get_device_properties(device_index)
        pci = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{pci}/local_cpulist") as fh:
            text = fh.read().strip()
        cpus = set()
        for part in text.split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except (OSError, ValueError, AttributeError, RuntimeError):
        return False


def all_reduce_null(null, group=None) -> None:
    """Sum a ``MoranNull`` over the ranks of ``group`` (counts travel as exact FP64 integers)."""
    _, ws = world(group)
    if ws > 1:
        packed = null.packed().contiguous()
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        null.unpack(packed)


def all_gather_columns(local, total_cols: int, device, group=None):
    """Gene-block sharding: every rank contributes ``local`` (numpy [rows, its block]); returns the
    concatenation over ranks in block order (numpy [rows, total_cols])."""
    import numpy as np

    rank, ws = world(group)
    if ws == 1:
        return local
    sizes = [block_slice(total_cols, r, ws) for r in range(ws)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros((local.shape[0], pad), dtype=torch.float64, device=device)
    buf[:, : local.shape[1]] = torch.from_numpy(np.ascontiguousarray(local)).to(device)
    out = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:, : hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=1).cpu().numpy()


def all_gather_column_blocks(local, sizes, device, group=None, max_rows: int = 1 << 20):
    """Every rank contributes the columns it owns of a per-cell matrix (numpy or device tensor ``[n, sizes[rank]]``,
    any dtype); returns the full ``[n, sum(sizes)]`` matrix in rank order on every rank.  The exchange goes
    through ``device`` in row chunks (NCCL all-gather on GPUs, gloo on CPU), so the staging buffers stay
    small next to the matrices themselves.  Ranks may own zero columns."""
    import numpy as np

    rank, ws = world(group)
    on_device = isinstance(local, torch.Tensor)  # columns already on `device`: no host round trip before the gather
    if ws == 1:
        return local.cpu().numpy() if on_device else local
    n = local.shape[0]
    pad = max(max(sizes), 1)
    if on_device:
        tdtype = local.dtype
        np_dtype = torch.empty(0, dtype=tdtype).numpy().dtype
    else:
        np_dtype = local.dtype
        tdtype = torch.from_numpy(np.empty(0, dtype=np_dtype)).dtype
    out = np.empty((n, int(sum(sizes))), dtype=np_dtype)
    starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    for r0 in range(0, n, max_rows):
        r1 = min(n, r0 + max_rows)
        buf = torch.zeros((r1 - r0, pad), dtype=tdtype, device=device)
        if sizes[rank] > 0:
            buf[:, : sizes[rank]] = local[r0:r1] if on_device else torch.from_numpy(np.ascontiguousarray(local[r0:r1])).to(device)
        parts = [torch.empty_like(buf) for _ in range(ws)]
        dist.all_gather(parts, buf, group=group)
        for r in range(ws):
            if sizes[r] > 0:
                out[r0:r1, starts[r]:starts[r + 1]] = parts[r][:, : sizes[r]].cpu().numpy()
    return out
