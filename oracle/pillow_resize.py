"""Integer restatement of Pillow's antialiased BILINEAR resize on uint8.  TEST INFRASTRUCTURE (oracle).

The reference pre-process is ``PIL.Image.fromarray(img).resize((320, 320), Image.BILINEAR)``
(perspective2d/perspectivefields.py:38-46).  The arithmetic lives in Pillow (third-party, unpinned in
requirements.txt:6; 12.2.0 in this image), ``src/libImaging/Resample.c``: ``precompute_coeffs``,
``normalize_coeffs_8bpc``, ``ImagingResampleHorizontal_8bpc`` then ``ImagingResampleVertical_8bpc``.
Published algorithm, per axis (in -> out):

    scale = in / out;  filterscale = max(scale, 1);  support = 1.0 * filterscale     (triangle filter)
    center = (x + 0.5) * scale
    xmin = max(0, int(center - support + 0.5));  xmax = min(in, int(center + support + 0.5))
    w_i  = max(0, 1 - |(i + xmin - center + 0.5) / filterscale|),  normalised by their sum   (double)
    k_i  = int(0.5 + w_i * 2**22)
    out  = clip8((sum_i px_i * k_i + 2**21) >> 22)

The horizontal pass runs first and is rounded to uint8 before the vertical pass; an axis whose size does
not change is skipped.  Pinned against Pillow itself in tests/test_oracle_pillow.py (Pillow is present on
both the build container and the GPU box).
"""
import numpy as np

PRECISION_BITS = 32 - 8 - 2


def precompute_coeffs(in_size, out_size):
    """Returns (bounds[out,2] int32 = (xmin, count), coeffs[out,ksize] int32 fixed-point, ksize)."""
    scale = float(in_size) / float(out_size)
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size)
        n = xmax - xmin
        ww = 0.0
        for x in range(n):
            a = (x + xmin - center + 0.5) * ss
            w = 1.0 - abs(a) if abs(a) < 1.0 else 0.0
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            kk[xx, :n] /= ww
        bounds[xx] = (xmin, n)
    coeffs = np.where(kk < 0, (-0.5 + kk * (1 << PRECISION_BITS)).astype(np.int64),
                      (0.5 + kk * (1 << PRECISION_BITS)).astype(np.int64)).astype(np.int32)
    return bounds, coeffs, ksize


def _resample_axis0(img, out_size):
    """Resample along axis 0 of a [in, ..., C] uint8 array."""
    in_size = img.shape[0]
    bounds, coeffs, ksize = precompute_coeffs(in_size, out_size)
    src = img.astype(np.int64)
    out = np.empty((out_size,) + img.shape[1:], np.uint8)
    for xx in range(out_size):
        xmin, n = bounds[xx]
        k = coeffs[xx, :n].astype(np.int64)
        acc = np.tensordot(k, src[xmin:xmin + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bilinear_u8(img, out_h, out_w):
    """``np.asarray(Image.fromarray(img).resize((out_w, out_h), Image.BILINEAR))`` for uint8 [H,W,C]."""
    assert img.dtype == np.uint8 and img.ndim == 3
    h, w = img.shape[:2]
    out = img
    if w != out_w:  # horizontal pass first
        out = np.swapaxes(_resample_axis0(np.swapaxes(out, 0, 1), out_w), 0, 1)
    if h != out_h:
        out = _resample_axis0(out, out_h)
    return np.ascontiguousarray(out)
