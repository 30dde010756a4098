"""TEST INFRASTRUCTURE ONLY — CPU restatement of overlap-tile inference (BASELINE config 4).

PARITY UNPINNED: the reference ships no tiling code — only result pictures
(``predictions_output_overlap_tile/*.png``) and prose (``README.md:104-106``,
``images/old readme unet.txt:73-87``, which names a non-existent ``inference_tiled.py``). The
semantics below are the U-Net paper's (Fig. 2) as this repository defines them:

  * the image is extended by mirroring (numpy ``reflect``) by the network margin, so that the valid
    convolutions' output covers every pixel of the image;
  * input tiles of size S ≡ 12 (mod 16) at output origins that are multiples of 16; each tile is one
    eval-mode forward of the oracle network (``unet_ref.unet_forward``, itself pinned against the
    reference's ``UNet``); its (S - 2·margin)² logits are written at the tile's origin, later
    tiles overwrite earlier ones where they overlap;
  * mask = ``softmax(logits)[1] > 0.5`` ⇔ ``z1 > z0`` → 255 (``scripts/predict.py:85-92``).

What stands in for golden vectors is an invariant (checked in tests/): under the two
alignment conditions every max-pool window of a tile coincides with one of the whole-image forward,
so the stitched logits equal ONE forward over the whole mirror-extended image.

Written independently of ``unet_segmentation_b200/tiling.py`` (numpy padding and plain loops
instead of gather indices) so that the product's host logic can be checked against it.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from . import unet_ref


def margin(levels: int = 5) -> int:
    """(input - output) / 2 of the network for a size without pooling floors."""
    size = 12 + 16 * 2 ** levels
    return (size - unet_ref.out_size(size, levels)) // 2


def tile_origins(h: int, w: int, tile_in: int, levels: int = 5) -> Tuple[int, int, List[Tuple[int, int]]]:
    """(tile_out, stride, [(y, x), ...]) in row-major order; stride = tile_out rounded down to a
    multiple of 16; tiles are added until they cover [0, h) x [0, w)."""
    assert tile_in % 16 == 12, "tile_in must be ≡ 12 (mod 16)"
    tile_out = tile_in - 2 * margin(levels)
    assert tile_out >= 16
    stride = tile_out - tile_out % 16

    def axis(n):
        out = [0]
        while out[-1] + tile_out < n:
            out.append(out[-1] + stride)
        return out

    return tile_out, stride, [(y, x) for y in axis(h) for x in axis(w)]


def mirror_extend(image: np.ndarray, tile_in: int, levels: int = 5) -> np.ndarray:
    """The image padded by the margin on the top / left and by margin + (covered - size) on the
    bottom / right, 'reflect' (no edge repeat)."""
    h, w = image.shape
    m = margin(levels)
    tile_out, _, origins = tile_origins(h, w, tile_in, levels)
    cover_h = max(y for y, _ in origins) + tile_out
    cover_w = max(x for _, x in origins) + tile_out
    return np.pad(image, ((m, m + cover_h - h), (m, m + cover_w - w)), mode="reflect")


@torch.no_grad()
def overlap_tile_logits(sd, image: torch.Tensor, tile_in: int, levels: int = 5) -> torch.Tensor:
    """image (H, W) fp32 -> stitched logits (n_classes, H, W), one oracle forward per tile."""
    h, w = image.shape
    dev = image.device
    tile_out, _, origins = tile_origins(h, w, tile_in, levels)
    ext = torch.from_numpy(mirror_extend(image.cpu().numpy(), tile_in, levels)).to(dev)
    out = None
    for (y, x) in origins:
        tile = ext[y:y + tile_in, x:x + tile_in][None, None]
        z = unet_ref.unet_forward(sd, tile, training=False, levels=levels)[0]
        assert z.shape[-2:] == (tile_out, tile_out)
        if out is None:
            out = torch.zeros(z.shape[0], h, w, dtype=z.dtype, device=dev)
        hh, ww = min(tile_out, h - y), min(tile_out, w - x)
        out[:, y:y + hh, x:x + ww] = z[:, :hh, :ww]
    return out


@torch.no_grad()
def whole_image_logits(sd, image: torch.Tensor, tile_in: int, levels: int = 5) -> torch.Tensor:
    """ONE oracle forward over the whole mirror-extended image, cropped to (H, W)."""
    h, w = image.shape
    ext = torch.from_numpy(mirror_extend(image.cpu().numpy(), tile_in, levels)).to(image.device)
    return unet_ref.unet_forward(sd, ext[None, None], training=False, levels=levels)[0, :, :h, :w]


def mask_from_logits(logits: torch.Tensor) -> torch.Tensor:
    """scripts/predict.py:85-92: (softmax[1] > 0.5) * 255 as uint8."""
    return (logits[1] > logits[0]).to(torch.uint8) * 255
