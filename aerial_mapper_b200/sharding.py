"""Column-stripe sharding of the map across ranks — host-side logic only.

Every grid cell is independent in both stages of the path, so the map shards with no data-path collective:
rank r owns the contiguous column stripe [c0, c1) of the column-major layers (a contiguous slab of rows*(c1-c0)
floats per layer) and computes it from the shared inputs.  One all-gather per step of the finished stripes gives
every rank the full map.  The functions below are backend-agnostic (torch tensors: NCCL on GPUs, gloo on CPU).
"""
import numpy as np


def stripe_width(cols, world):
    """Equal stripe width for a single equal-count all-gather (the last stripes may be partly or wholly padding)."""
    return (cols + world - 1) // world


def stripe_range(cols, rank, world):
    """Columns [c0, c1) owned by `rank` (c0 == c1 for a rank that owns nothing)."""
    w = stripe_width(cols, world)
    c0 = min(rank * w, cols)
    c1 = min(c0 + w, cols)
    return c0, c1


def pack_slabs(torch, slabs, rows, width, out):
    """Copy each layer's slab (1-D tensor of rows*(c1-c0) floats) into `out` [n_layers, rows*width] (zero padded)."""
    for k, t in enumerate(slabs):
        n = t.numel()
        out[k, :n].copy_(t)
        if n < rows * width:
            out[k, n:].zero_()
    return out


def all_gather_stripes(torch, dist, packed, world):
    """One collective: gathered[r] = rank r's packed slabs."""
    gathered = torch.empty((world,) + tuple(packed.shape), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(gathered.view(-1), packed.view(-1))
    return gathered


def unpack_full(gathered, rows, cols, world, n_layers):
    """gathered [world, n_layers, rows*width] -> list of n_layers numpy float32 arrays (rows, cols), F order."""
    width = stripe_width(cols, world)
    g = gathered.cpu().numpy().reshape(world, n_layers, width, rows)  # slab memory is column-major: [col][row]
    out = []
    for k in range(n_layers):
        full = np.empty((rows, cols), dtype=np.float32, order="F")
        for r in range(world):
            c0, c1 = stripe_range(cols, r, world)
            if c1 > c0:
                full[:, c0:c1] = g[r, k, :c1 - c0, :].T
        out.append(full)
    return out
