"""Screen-space strip partition for multi-GPU rendering (new work, the reference has
no multi-GPU path). Rank g of G owns whole tile rows [y0, y1) so its output is one contiguous block
of the row-major RGBA8 frame; the frame is assembled inside the library (PFCudaRendererGatherFrame).
The same arithmetic as PFCudaStripOfRank (csrc/renderer.cu strip_of_rank), restated for host code that
has no GPU library loaded (tests/test_strip_gloo.py checks the two against each other)."""
from __future__ import annotations

TILE = 16


def tile_rows(height_px: int) -> int:
    return (height_px + TILE - 1) // TILE


def strip_rows(height_px: int, world: int, rank: int) -> tuple[int, int]:
    """Tile rows [y0, y1) of `rank`: as equal as whole rows allow (rows * rank // world)."""
    rows = tile_rows(height_px)
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} of {world}")
    if rows < world:
        raise ValueError(f"{rows} tile rows cannot be split across {world} ranks")
    return rows * rank // world, rows * (rank + 1) // world


def strip_pixel_rows(height_px: int, world: int, rank: int) -> tuple[int, int]:
    y0, y1 = strip_rows(height_px, world, rank)
    return min(y0 * TILE, height_px), min(y1 * TILE, height_px)
