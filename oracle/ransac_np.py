"""numpy restatement of the RANSAC scoring loops of the match list's consumer -- TEST INFRASTRUCTURE ONLY.

  CameraPoseEstimation.GetFundamentalMatrix   dotnet_src/ImageProcessing/CameraPoseEstimation.cs:26-94

The arithmetic lives in a third-party dependency that is not vendored in the reference tree:
MathNet.Numerics 5.0.0 (dotnet_src/ImageProcessing/ImageProcessing.csproj), single precision.  Its managed
provider evaluates ``F.Multiply(v)`` and ``DotProduct`` as plain left-to-right accumulations of products (no
fused multiply-add under the .NET JIT); that published behaviour is what is restated here.  There is no golden
vector for this step in the reference (the caller is commented out, Program.cs:207-249; the sampler is an
unseeded ``new Random()``, :35) -- parity unpinned.
"""
from __future__ import annotations

import numpy as np


def residuals(F: np.ndarray, xy1: np.ndarray, xy2: np.ndarray) -> np.ndarray:
    """float32[n_hyp, n]: (F . (x2, y2, 1)) . (x1, y1, 1)  (CameraPoseEstimation.cs:66-71)."""
    F = np.asarray(F, dtype=np.float32).reshape(-1, 3, 3)
    x1 = np.asarray(xy1)[:, 0].astype(np.float32)[None, :]
    y1 = np.asarray(xy1)[:, 1].astype(np.float32)[None, :]
    x2 = np.asarray(xy2)[:, 0].astype(np.float32)[None, :]
    y2 = np.asarray(xy2)[:, 1].astype(np.float32)[None, :]
    r = []
    for i in range(3):
        a, b, c = (F[:, i, k][:, None] for k in range(3))
        r.append((a * x2 + b * y2) + c)                  # float32 products and sums, left to right
    return ((r[0] * x1 + r[1] * y1) + r[2]).astype(np.float32)


def score(F, valid, xy1, xy2, threshold):
    """(counts int32[n_hyp] with -1 for skipped hypotheses, best index or -1, inlier mask of the best)."""
    res = residuals(F, xy1, xy2)
    inl = res <= np.float32(threshold)                   # :73, signed
    counts = inl.sum(axis=1).astype(np.int32)
    if valid is not None:
        counts = np.where(np.asarray(valid).astype(bool), counts, -1).astype(np.int32)
    best, best_count = -1, 0
    for k, c in enumerate(counts.tolist()):              # :79-84: strictly more inliers than every earlier sample
        if c > best_count:
            best, best_count = k, c
    mask = inl[best].astype(np.uint8) if best >= 0 else np.zeros(len(np.asarray(xy1)), dtype=np.uint8)
    return counts, best, mask
