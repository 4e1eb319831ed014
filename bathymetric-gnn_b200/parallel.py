"""Multi-GPU sharding of a survey's tile list: one process per GPU, NO data-path
collective.  Overlapping tiles are independent units; each rank
owns a contiguous band of tile-rows and private accumulators for the rows its
tiles touch; only the per-band partial planes are gathered (once, at the end).

Seams between bands (rows covered by tiles of two ranks): the discrete channel
is an arg-max by (confidence, earlier tile wins ties) and merges bit-exactly;
the continuous channels merge as partial sums `(A+B) + (C+D)` where the
single-GPU order is `((A+B)+C)+D`, i.e. <= 2 ulp32 on seam cells only (stated
tolerance; everything off the seams is bit-identical to the 1-GPU run).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

from .tiling import TileSpec


def band_partition(num_tile_rows: int, world: int) -> List[Tuple[int, int]]:
    """Balanced contiguous [lo, hi) ranges of tile-rows, one per rank (empty when world > rows)."""
    base, extra = divmod(num_tile_rows, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def band_specs(specs: Sequence[TileSpec], band: Tuple[int, int]) -> List[TileSpec]:
    return [s for s in specs if band[0] <= s.tile_row < band[1]]


def band_rows(specs: Sequence[TileSpec]) -> Tuple[int, int]:
    """Row range [r0, r1) of the survey touched by a band's tiles."""
    if not specs:
        return (0, 0)
    return min(s.row_start for s in specs), max(s.row_end for s in specs)


def shift_specs(specs: Sequence[TileSpec], row0: int) -> List[TileSpec]:
    return [TileSpec(s.row_start - row0, s.col_start, s.row_end - row0, s.col_end, s.tile_row, s.tile_col) for s in specs]


def merge_band_partials(shape: Tuple[int, int], parts: Sequence[Dict[str, torch.Tensor]],
                        rows: Sequence[Tuple[int, int]]) -> Dict[str, torch.Tensor]:
    """Combine UNFINALISED per-band planes (in band order = reference tile order) into full-survey planes.
    Each part holds: acc_<ch>, wsum_<ch> for continuous channels, `classification`, `tracker`."""
    dev = parts[0]["classification"].device
    full: Dict[str, torch.Tensor] = {}
    for k in parts[0]:
        if k.startswith("wsum_"):
            full[k] = torch.zeros(shape, dtype=torch.float32, device=dev)
        elif k == "tracker":
            full[k] = torch.full(shape, -1.0, dtype=torch.float32, device=dev)
        else:
            full[k] = torch.full(shape, float("nan"), dtype=torch.float32, device=dev)
    for p, (r0, r1) in zip(parts, rows):
        if r1 <= r0:
            continue
        sl = slice(r0, r1)
        for k, v in p.items():
            v = v.to(dev)
            if k.startswith("acc_"):
                cur = full[k][sl]
                both = ~torch.isnan(cur) & ~torch.isnan(v)
                full[k][sl] = torch.where(both, cur + v, torch.where(torch.isnan(cur), v, cur))
            elif k.startswith("wsum_"):
                full[k][sl] += v
        # discrete: later band wins only with strictly higher confidence (tiling.py:413-417)
        cur_c, cur_t = full["classification"][sl], full["tracker"][sl]
        new_c, new_t = p["classification"].to(dev), p["tracker"].to(dev)
        take = ~torch.isnan(new_c) & ((new_t > cur_t) | torch.isnan(cur_c))
        full["classification"][sl] = torch.where(take, new_c, cur_c)
        full["tracker"][sl] = torch.where(take, new_t, cur_t)
    return full


def finalize_merged(full: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """tiling.py:289-292 on the merged planes."""
    out = {"classification": full["classification"]}
    for k, v in full.items():
        if k.startswith("acc_"):
            w = full["wsum_" + k[4:]]
            out[k[4:]] = torch.where(w > 0, v / w, v)
    return out
