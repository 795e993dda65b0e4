"""Scenario ensembles: thousands of members that share one network and differ in forcing.

Members are independent (each is its own `Model` in the reference, build/libribasim.jl:6), so
the member axis shards over the GPUs of one box with no exchange inside the time loop; the only
collective is the final gather of the results.
"""

from __future__ import annotations

import numpy as np

PERTURBED = ("precipitation", "potential_evaporation", "drainage")


def member_range(n_members: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous block of the member axis owned by `rank`: [lo, hi)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, rem = divmod(n_members, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def perturbation_factors(member: int, day: int, n_basin: int, n_fb: int, sigma: float = 0.2):
    """iid lognormal(0, sigma) factors of member `member` on day `day` for the perturbed basin
    forcings and the boundary inflows (BASELINE.json config 4).  The stream
    depends on (member, day) only, never on the shard layout."""
    rng = np.random.default_rng([int(member), int(day)])
    return rng.lognormal(0.0, sigma, size=(len(PERTURBED), n_basin)), \
        rng.lognormal(0.0, sigma, size=n_fb)


def perturbed_forcing(base: dict[str, np.ndarray], fb_base: np.ndarray, members: range, day: int,
                      out: dict[str, np.ndarray] | None = None, sigma: float = 0.2):
    """Per-member forcing arrays `[n_entity][n_members_local]` for one day.

    `base[name]` is the unperturbed `[n_basin]` vertical flux, `fb_base` the `[n_fb]` boundary
    inflow.  `out` may hold preallocated (pinned) arrays to fill in place."""
    n_basin = len(next(iter(base.values())))
    n_fb = len(fb_base)
    nm = len(members)
    if out is None:
        out = {k: np.empty((n_basin, nm)) for k in PERTURBED}
        out["fb_rate"] = np.empty((n_fb, nm))
    for j, m in enumerate(members):
        f, g = perturbation_factors(m, day, n_basin, n_fb, sigma)
        for i, k in enumerate(PERTURBED):
            out[k][:, j] = base[k] * f[i]
        out["fb_rate"][:, j] = fb_base * g
    return out


def gather_members(local: np.ndarray, n_members: int, group=None) -> np.ndarray | None:
    """Final result gather: `local` is `[n_entity][n_local]` on every rank; rank 0 gets
    `[n_entity][n_members]` in member order.  Uses torch.distributed (NCCL for CUDA tensors,
    gloo for CPU tensors); with a single process it is the identity."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return np.asarray(local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    t = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local))
    n_entity = t.shape[0]
    widths = [member_range(n_members, r, world) for r in range(world)]
    wmax = max(hi - lo for lo, hi in widths)
    pad = torch.zeros((n_entity, wmax), dtype=t.dtype, device=t.device)
    pad[:, : t.shape[1]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    if rank != 0:
        return None
    out = torch.cat([b[:, : hi - lo] for b, (lo, hi) in zip(bufs, widths)], dim=1)
    return out.cpu().numpy()
