"""Seeded synthetic screenshot pairs (NumPy only; no cv2 so it runs anywhere).

Workload definitions follow SURVEY.md section 8(d): a T (texture) and an S (screenshot-like) generator, a known
sub-pixel translation between expected and target, and an optional "defect" (one rectangle moved by 8 px)
that must flip the status to SUSPICIOUS.  The translation is applied with a separable 4-tap cubic
convolution (a = -0.75), which for a pure translation is exactly bicubic interpolation.
"""
from __future__ import annotations

import numpy as np

SHIFT = (0.37, -0.61)  # target(x, y) = canvas(x + 0.37, y - 0.61)  => true flow (dx, dy) = (-0.37, +0.61)
_PAD = 32


def _gauss_blur(a: np.ndarray, sigma: float) -> np.ndarray:
    r = int(np.ceil(4 * sigma))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-x * x / (2 * sigma * sigma))
    k = (k / k.sum()).astype(np.float32)
    p = np.pad(a, ((0, 0), (r, r)), mode="reflect")
    out = np.zeros_like(a)
    for i, kv in enumerate(k):
        out += kv * p[:, i:i + a.shape[1]]
    p = np.pad(out, ((r, r), (0, 0)), mode="reflect")
    out2 = np.zeros_like(a)
    for i, kv in enumerate(k):
        out2 += kv * p[i:i + a.shape[0], :]
    return out2


def _cubic_weights(f: float, a: float = -0.75) -> np.ndarray:
    def w(t):
        t = abs(t)
        if t <= 1:
            return (a + 2) * t ** 3 - (a + 3) * t ** 2 + 1
        if t < 2:
            return a * t ** 3 - 5 * a * t ** 2 + 8 * a * t - 4 * a
        return 0.0
    return np.array([w(1 + f), w(f), w(1 - f), w(2 - f)], np.float64)


def _shift_frac(canvas: np.ndarray, sx: float, sy: float) -> np.ndarray:
    """out(x, y) = canvas(x + sx, y + sy), bicubic, reflect border."""
    def shift_axis(a, s, axis):
        i = int(np.floor(s))
        f = s - i
        wts = _cubic_weights(f)
        pad = [(0, 0), (0, 0)]
        pad[axis] = (8, 8)
        p = np.pad(a, pad, mode="reflect")
        out = np.zeros_like(a, dtype=np.float64)
        n = a.shape[axis]
        for t in range(4):
            off = 8 + i - 1 + t
            sl = [slice(None), slice(None)]
            sl[axis] = slice(off, off + n)
            out += wts[t] * p[tuple(sl)]
        return out
    return shift_axis(shift_axis(canvas.astype(np.float64), sx, 1), sy, 0)


def _to_u8(a: np.ndarray) -> np.ndarray:
    return np.clip(np.rint(a), 0, 255).astype(np.uint8)


def canvas_texture(W: int, H: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    a = rng.random((H + 2 * _PAD, W + 2 * _PAD), dtype=np.float32)
    a = _gauss_blur(a, 2.0)
    a = (a - a.min()) / (a.max() - a.min()) * 255.0
    return a.astype(np.float32)


def _rects(rng, W, H, n):
    out = []
    for _ in range(n):
        x = int(rng.integers(0, W)); y = int(rng.integers(0, H))
        w = int(rng.integers(8, 300)); h = int(rng.integers(4, 60))
        g = int(rng.integers(0, 255))
        out.append((x, y, w, h, g))
    return out


def canvas_screenshot(W: int, H: int, seed: int, defect: bool = False):
    """Returns (canvas, canvas_with_defect_or_None).  400 filled rectangles + 4000 dark strokes on white."""
    rng = np.random.default_rng(seed)
    CW, CH = W + 2 * _PAD, H + 2 * _PAD
    scale = (W * H) / (1920 * 1080)
    rects = _rects(rng, CW, CH, max(20, int(400 * scale)))
    strokes = []
    for _ in range(max(200, int(4000 * scale))):
        x = int(rng.integers(0, CW)); y = int(rng.integers(0, CH))
        h = int(rng.integers(1, 3)); w = int(rng.integers(2, 12))
        g = int(rng.integers(0, 120))
        strokes.append((x, y, w, h, g))

    def paint(rl):
        c = np.full((CH, CW), 255.0, np.float32)
        for (x, y, w, h, g) in rl:
            c[y:y + h, x:x + w] = g
        for (x, y, w, h, g) in strokes:
            c[y:y + h, x:x + w] = g
        return c

    base = paint(rects)
    moved = None
    if defect:
        # move the last-painted sizeable rectangle that lies well inside the frame by 8 px
        rl = list(rects)
        for i in range(len(rl) - 1, -1, -1):
            x, y, w, h, g = rl[i]
            if w >= 60 and h >= 24 and _PAD + 40 < x < CW - w - 80 and _PAD + 40 < y < CH - h - 80 and g < 200:
                rl[i] = (x + 8, y + 8, w, h, g)
                break
        moved = paint(rl)
    return base, moved


def make_pair(kind: str, W: int, H: int, seed: int, defect: bool = False):
    """kind 'T' or 'S'.  Returns (expected u8 HxW, target u8 HxW)."""
    if kind == "T":
        canvas = canvas_texture(W, H, seed)
        tgt_canvas = canvas
    elif kind == "S":
        canvas, moved = canvas_screenshot(W, H, seed, defect)
        tgt_canvas = moved if defect else canvas
    else:
        raise ValueError(kind)
    expected = _to_u8(canvas[_PAD:_PAD + H, _PAD:_PAD + W])
    shifted = _shift_frac(tgt_canvas, SHIFT[0], SHIFT[1])
    target = _to_u8(shifted[_PAD:_PAD + H, _PAD:_PAD + W])
    return expected, target


def pool_pairs(n: int, W: int = 1920, H: int = 1080, seed0: int = 100):
    """Config 5's pool: seeds seed0.., alternating S/T, every 8th with a defect."""
    out = []
    for i in range(n):
        kind = "S" if i % 2 == 0 else "T"
        out.append(make_pair(kind, W, H, seed0 + i, defect=(kind == "S" and i % 8 == 0)))
    return out
