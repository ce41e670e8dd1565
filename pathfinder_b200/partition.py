"""Screen-space strip partition for multi-GPU rendering (new work, the reference has
no multi-GPU path). Rank g of G owns whole tile rows [y0, y1) so its output is one contiguous block
of the row-major RGBA8 frame; the frame is assembled with one all-gather."""
from __future__ import annotations

TILE = 16


def tile_rows(height_px: int) -> int:
    return (height_px + TILE - 1) // TILE


def strip_rows(height_px: int, world: int, rank: int) -> tuple[int, int]:
    """Equal strips of whole tile rows; requires world | tile_rows so that the gather is uniform
    (torch.distributed.all_gather_into_tensor)."""
    rows = tile_rows(height_px)
    if rows % world != 0:
        raise ValueError(f"{rows} tile rows do not divide evenly across {world} ranks")
    per = rows // world
    return rank * per, (rank + 1) * per


def strip_pixel_rows(height_px: int, world: int, rank: int) -> tuple[int, int]:
    y0, y1 = strip_rows(height_px, world, rank)
    return y0 * TILE, min(y1 * TILE, height_px)
