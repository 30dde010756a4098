"""Device-side input pipeline.

The reference builds every sample on the host (``utils/dataset.py:92-110``: ``ToTensor`` →
float32 in [0, 1], ``(mask > 0).long()``, ``weight_map.float()``) and crops target / weight map to the
logits size on the device (``scripts/train.py:39-51,118-126``). Here the host ships the batch as it
lies on disk — uint8 frames, uint8/uint16 instance masks, float32/float64 weight maps — and ONE
kernel (``ub_prepare_batch``) produces the three tensors the training step consumes:

    images  float32 (N, 1, H, W)   = u8 / 255
    targets int64   (N, h, w)      = (label > 0), centre-cropped
    weights float32 (N, h, w)      = float(weight map), centre-cropped

bit-exactly what the reference pipeline yields, with 6–11 bytes per pixel crossing PCIe instead of
16. ``DeviceBatchPreparer`` double-buffers the host→device copies on a side stream.

Row N4: ``weight_maps_from_labels`` computes the weight maps themselves on the device
(``ub_weight_map`` = ``calculate_weight_map`` of ``scripts/preprocess_data.py:17-77``, bit-exact
against the reference's stored ``.npy`` maps), so the host ships 3 bytes per pixel (uint8 frame +
uint16 labels) and ``DeviceBatchPreparer.submit(images, labels, None)`` needs no stored maps;
``elastic_deform`` is the reference's training augmentation (``utils/augmentations.py:4-39``,
``elastic_deform_image_and_mask``; enabled in ``scripts/train.py:34-36`` with alpha 2000, sigma 20)
for a whole batch in three kernels (``ub_elastic_deform``), bit-exact against scipy given the same
uniform draws.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check


def _p(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(t.data_ptr() if t is not None else 0)


def prepare_batch(images_u8: torch.Tensor, labels: Optional[torch.Tensor],
                  weight_maps: Optional[torch.Tensor], out_hw: Tuple[int, int]):
    """images_u8 (N, H, W) uint8, labels (N, H, W) uint8|uint16|int16, weight_maps (N, H, W)
    float32|float64 — CUDA tensors. Returns (images, targets, weights) as described above."""
    if not images_u8.is_cuda or images_u8.dtype != torch.uint8 or images_u8.dim() != 3:
        raise ValueError("prepare_batch expects a CUDA uint8 tensor of shape (N, H, W); the B200 "
                         "input pipeline has no CPU path")
    n, h, w = images_u8.shape
    oh, ow = out_hw
    dev = images_u8.device
    images_u8 = images_u8.contiguous()
    lb = 0
    if labels is not None:
        if labels.shape != images_u8.shape or labels.dtype not in (torch.uint8, torch.uint16, torch.int16):
            raise ValueError("labels must be (N, H, W) uint8 / uint16")
        labels = labels.contiguous()
        lb = labels.element_size()
    wb = 0
    if weight_maps is not None:
        if weight_maps.shape != images_u8.shape or weight_maps.dtype not in (torch.float32, torch.float64):
            raise ValueError("weight_maps must be (N, H, W) float32 / float64")
        weight_maps = weight_maps.contiguous()
        wb = weight_maps.element_size()
    image = torch.empty(n, 1, h, w, dtype=torch.float32, device=dev)
    target = torch.empty(n, oh, ow, dtype=torch.int64, device=dev) if labels is not None else None
    weight = torch.empty(n, oh, ow, dtype=torch.float32, device=dev) if weight_maps is not None else None
    lib = _lib.load()
    with torch.cuda.device(dev):
        check(lib.ub_prepare_batch(_p(images_u8), _p(labels), lb, _p(weight_maps), wb, n, h, w, oh, ow,
                                   _p(image), _p(target), _p(weight),
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)),
              "ub_prepare_batch")
    return image, target, weight


def weight_maps_from_labels(labels: torch.Tensor, w0: float = 10, sigma: float = 5,
                            dtype: torch.dtype = torch.float64) -> torch.Tensor:
    """``calculate_weight_map(mask, w0, sigma)`` (reference scripts/preprocess_data.py:17-77, defaults
    W0 = 10, SIGMA = 5 of :14-15) for a batch: labels (N, H, W) or (H, W) uint8|uint16|int16 CUDA
    instance masks -> weight maps of the same shape, float64 as the reference stores them (or
    float32 = what ``utils/dataset.py:110`` hands to the loss). Bit-exact; see
    ``csrc/weight_map.cuh`` for why the reference's border term is the constant ``w0``."""
    if not labels.is_cuda or labels.dtype not in (torch.uint8, torch.uint16, torch.int16) \
            or labels.dim() not in (2, 3):
        raise ValueError("weight_maps_from_labels expects a CUDA uint8 / uint16 tensor of shape "
                         "(N, H, W) or (H, W); the B200 input pipeline has no CPU path")
    if dtype not in (torch.float32, torch.float64):
        raise ValueError("dtype must be torch.float32 or torch.float64")
    lab = labels.contiguous()
    if lab.dim() == 2:
        lab = lab.unsqueeze(0)
    n, h, w = lab.shape
    out = torch.empty(n, h, w, dtype=dtype, device=lab.device)
    if out.numel() == 0:
        return out.reshape(labels.shape)
    counts = torch.empty(n, dtype=torch.int32, device=lab.device)
    lib = _lib.load()
    with torch.cuda.device(lab.device):
        check(lib.ub_weight_map(_p(lab), lab.element_size(), n, h, w, float(w0), float(sigma), _p(out),
                                out.element_size(), _p(counts),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream)),
              "ub_weight_map")
    return out.reshape(labels.shape)


def gaussian_taps(sigma: float, truncate: float = 4.0) -> np.ndarray:
    """The 1-D kernel ``scipy.ndimage.gaussian_filter`` builds (``_gaussian_kernel1d``): radius
    int(truncate * sigma + 0.5), exp(-0.5 / sigma^2 * x^2) normalised by its sum. Evaluated with
    numpy on the host so that the weights carry numpy's ``exp`` rounding (2 r + 1 doubles per
    call — plumbing, not a compute path)."""
    sd = float(sigma)
    if not sd > 0.0:
        raise ValueError("sigma must be positive")
    radius = int(truncate * sd + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sd * sd) * x ** 2)
    return phi / phi.sum()


_TAPS_CACHE: dict = {}


def _device_taps(sigma: float, device: torch.device) -> Tuple[torch.Tensor, int]:
    """Gaussian taps as a device tensor, cached per (sigma, device): the upload is a synchronous
    pageable copy that should not sit in a per-batch loop."""
    key = (float(sigma), str(device))
    hit = _TAPS_CACHE.get(key)
    if hit is None:
        taps_np = gaussian_taps(sigma)
        hit = (torch.from_numpy(taps_np).to(device), (len(taps_np) - 1) // 2)
        _TAPS_CACHE[key] = hit
    return hit


def reference_noise(seeds, shape: Tuple[int, int]) -> torch.Tensor:
    """The uniform draws the reference makes for each sample (``np.random.RandomState(seed)``,
    ``rand(*shape)`` for dx then for dy; utils/augmentations.py:19-28, seeds from
    utils/dataset.py:84) as a pinned float64 host tensor (2, N, H, W) for ``elastic_deform(noise=)``.
    Use it to reproduce a reference run; leave ``noise=None`` to draw on the device instead."""
    out = np.empty((2, len(seeds)) + tuple(shape), np.float64)
    for k, seed in enumerate(seeds):
        rs = np.random.RandomState(seed)
        out[0, k] = rs.rand(*shape)
        out[1, k] = rs.rand(*shape)
    t = torch.from_numpy(out)
    return t.pin_memory() if torch.cuda.is_available() else t


def elastic_deform(images_u8: Optional[torch.Tensor], labels: Optional[torch.Tensor],
                   alpha: float = 2000, sigma: float = 20, noise: Optional[torch.Tensor] = None,
                   generator: Optional[torch.Generator] = None, labels_as_uint8: bool = False):
    """``elastic_deform_image_and_mask`` (reference utils/augmentations.py:4-39) for a batch:
    images_u8 (N, H, W) uint8 -> bilinear-resampled uint8; labels (N, H, W) uint8|uint16 ->
    nearest-neighbour-resampled labels (``labels_as_uint8`` wraps uint16 ids modulo 256 like the
    ``astype(np.uint8)`` of utils/dataset.py:93). ``noise``: (2, N, H, W) float64 CUDA tensor of
    uniform [0, 1) draws (``reference_noise`` reproduces the reference's seeds); None draws them on
    the device with ``generator``. Returns (images, labels); an absent input gives None."""
    ref = images_u8 if images_u8 is not None else labels
    if ref is None or not ref.is_cuda or ref.dim() != 3:
        raise ValueError("elastic_deform expects CUDA tensors of shape (N, H, W); the B200 input "
                         "pipeline has no CPU path")
    n, h, w = ref.shape
    dev = ref.device
    if images_u8 is not None and (images_u8.dtype != torch.uint8 or images_u8.shape != ref.shape):
        raise ValueError("images must be (N, H, W) uint8")
    if labels is not None and (labels.shape != ref.shape or labels.device != dev or
                               labels.dtype not in (torch.uint8, torch.uint16, torch.int16)):
        raise ValueError("labels must be (N, H, W) uint8 / uint16 on the images' device")
    if noise is None:
        noise = torch.rand(2, n, h, w, dtype=torch.float64, device=dev, generator=generator)
    if noise.shape != (2, n, h, w) or noise.dtype != torch.float64 or noise.device != dev:
        raise ValueError("noise must be a (2, N, H, W) float64 tensor on the images' device")
    taps, radius = _device_taps(sigma, dev)
    images_u8 = images_u8.contiguous() if images_u8 is not None else None
    labels = labels.contiguous() if labels is not None else None
    noise = noise.contiguous()
    img_out = torch.empty_like(images_u8) if images_u8 is not None else None
    lab_out = None
    if labels is not None:
        lab_out = torch.empty(n, h, w, dtype=torch.uint8 if labels_as_uint8 else labels.dtype, device=dev)
    if n * h * w == 0:
        return img_out, lab_out
    lib = _lib.load()
    ws = torch.empty(int(lib.ub_elastic_workspace_bytes(n, h, w)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.ub_elastic_deform(_p(images_u8), _p(labels), labels.element_size() if labels is not None else 0,
                                    n, h, w, _p(noise), _p(taps), radius, float(alpha), _p(img_out),
                                    _p(lab_out), lab_out.element_size() if lab_out is not None else 0,
                                    _p(ws), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
              "ub_elastic_deform")
    return img_out, lab_out


class DeviceBatchPreparer:
    """Double-buffered host→device staging + the device-side sample pipeline on a side stream:
    ``submit`` enqueues, for the NEXT batch, the compact copies and every kernel of the pipeline
    (weight maps, elastic deformation, ``prepare_batch``) on the preparer's own stream, so that they
    overlap the training step running on the compute stream; ``get`` only makes the compute stream
    wait for the finished tensors."""

    def __init__(self, device, out_hw: Tuple[int, int], w0: float = 10, sigma: float = 5,
                 augment: Optional[Tuple[float, float]] = None,
                 generator: Optional[torch.Generator] = None):
        """``augment=(alpha, sigma)`` applies the reference's elastic deformation to frame and
        labels on the device (utils/dataset.py:83-94; (2000, 20) in scripts/train.py:35-36), with
        uniform draws from ``generator``. As in the reference, the weight map is NOT deformed: it
        is the stored map, or — when none is submitted — the map of the undeformed labels."""
        self.device = torch.device(device)
        self.out_hw = out_hw
        self.w0, self.sigma = w0, sigma   # used when a batch is submitted without stored maps
        self.augment, self.generator = augment, generator
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [None, None]
        self.k = 0

    def _process(self, images, labels, wmaps):
        if wmaps is None and labels is not None:
            wmaps = weight_maps_from_labels(labels, self.w0, self.sigma, torch.float32)
        if self.augment is not None:
            # the reference casts the deformed mask to uint8 (utils/dataset.py:93)
            images, labels = elastic_deform(images, labels, self.augment[0], self.augment[1],
                                            generator=self.generator, labels_as_uint8=True)
        return prepare_batch(images, labels, wmaps, self.out_hw)

    def submit(self, images_u8, labels, weight_maps):
        """Pinned host tensors (uint8 images, uint8/uint16 labels, float32/float64 weight maps).
        ``weight_maps=None`` with labels present: the maps are computed on the device (row N4).
        Returns a slot index for ``get``; at most two batches may be in flight."""
        k = self.k
        self.k ^= 1
        with torch.cuda.stream(self.stream):
            # staging copies and intermediates are allocated, used and freed on this stream only
            dev = [t.to(self.device, non_blocking=True) if t is not None else None
                   for t in (images_u8, labels, weight_maps)]
            out = self._process(*dev)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.slots[k] = (out, ev)
        return k

    def get(self, k):
        """(images fp32 (N,1,H,W), targets int64 (N,h,w), weights fp32 (N,h,w)) of slot ``k``, ordered
        after the preparer's stream on the current stream."""
        out, ev = self.slots[k]
        self.slots[k] = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in out:
            if t is not None:
                t.record_stream(cur)   # produced on the side stream, consumed (and freed) on this one
        return out
