"""Host mirror of the match list's consumer, the fundamental-matrix RANSAC.

  CameraPoseEstimation.GetFundamentalMatrix      dotnet_src/ImageProcessing/CameraPoseEstimation.cs:26-94
  CameraPoseEstimation.EstimateFundamentalMatrix :204-250, CalculateTransformationMatrix :252-275

The hot loops -- every hypothesis against every pair (:53-84) -- run in the CUDA kernels behind
``pgm_ransac_score``.  The per-sample 8-point estimates are a batched SVD of [numPairsPerSample x 9] systems, done
on the same GPU with ``torch.linalg`` (library plumbing; upstream uses MathNet.Numerics' SVD).  Upstream quirks kept:
``Math.Pow(2 / msd, 1 / 2)`` is ``Pow(x, 0) == 1`` (integer division, :265), so the normalisation only translates;
the estimate is not forced to rank 2, and samples whose matrix does not have rank 2 are skipped (:46-51).
The sampler upstream is an unseeded ``new Random()`` (:35): here it is seeded, so runs are reproducible; results
can only be compared with the reference statistically (parity unpinned).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .keypoint import KeypointPair
from .keypoint_matching import Matcher, default_matcher


class InvalidOperationException(ValueError):
    """System.InvalidOperationException (CameraPoseEstimation.cs:28-32, :206-207)."""


def _coords(pairs: Sequence[KeypointPair]) -> Tuple[np.ndarray, np.ndarray]:
    xy1 = np.array([[p.Keypoint1.Coordinate.X, p.Keypoint1.Coordinate.Y] for p in pairs], dtype=np.int32).reshape(-1, 2)
    xy2 = np.array([[p.Keypoint2.Coordinate.X, p.Keypoint2.Coordinate.Y] for p in pairs], dtype=np.int32).reshape(-1, 2)
    return xy1, xy2


class CameraPoseEstimation:
    def __init__(self, device: int = 0, matcher: Optional[Matcher] = None, seed: int = 0, require_rank2: bool = True):
        self._matcher = matcher or default_matcher(device)
        self._seed = seed
        self._require_rank2 = require_rank2
        self.last_hypotheses = None          # (F float32[S, 3, 3], valid uint8[S], counts int32[S]) of the last call

    # -- :204-250, batched over samples ---------------------------------------------------------
    def estimate_fundamental_matrices(self, xy1: np.ndarray, xy2: np.ndarray):
        """xy1, xy2: int[S, P, 2] sampled coordinates.  Returns (F float32[S, 3, 3], rank int[S])."""
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("CameraPoseEstimation needs a CUDA device (there is no CPU path)")
        dev = torch.device("cuda", self._matcher.device)
        a = torch.from_numpy(np.ascontiguousarray(xy1)).to(dev)
        b = torch.from_numpy(np.ascontiguousarray(xy2)).to(dev)
        S, P, _ = a.shape
        if P < 8:
            raise InvalidOperationException("At least 8 keypoint pairs must be provided")

        def transform(c):                      # CalculateTransformationMatrix: centroid in double, scale == 1
            cen = c.to(torch.float64).mean(dim=1)
            T = torch.eye(3, dtype=torch.float32, device=dev).repeat(S, 1, 1)
            T[:, 0, 2] = -cen[:, 0].to(torch.float32)
            T[:, 1, 2] = -cen[:, 1].to(torch.float32)
            return T
        T1, T2 = transform(a), transform(b)
        ones = torch.ones((S, P, 1), dtype=torch.float32, device=dev)
        c1 = torch.cat([a.to(torch.float32), ones], dim=2) @ T1.transpose(1, 2)      # T . (x, y, 1)
        c2 = torch.cat([b.to(torch.float32), ones], dim=2) @ T2.transpose(1, 2)
        mat = torch.stack([c1[..., 0] * c2[..., 0], c1[..., 0] * c2[..., 1], c1[..., 0],
                           c1[..., 1] * c2[..., 0], c1[..., 1] * c2[..., 1], c1[..., 1],
                           c2[..., 0], c2[..., 1], torch.ones_like(c1[..., 0])], dim=2)
        vh = torch.linalg.svd(mat, full_matrices=True).Vh
        last = vh[:, -1, :]
        F0 = last.reshape(S, 3, 3).transpose(1, 2)            # DenseOfColumnMajor(3, 3, lastRow)
        F = T2.transpose(1, 2) @ F0 @ T1
        sv = torch.linalg.svdvals(F)
        tol = torch.from_numpy(np.spacing(sv[:, 0].cpu().numpy().astype(np.float32))).to(dev) * 3.0   # Svd.Rank
        rank = (sv > tol[:, None]).sum(dim=1)
        return F.contiguous().cpu().numpy(), rank.cpu().numpy()

    def EstimateFundamentalMatrix(self, keypointPairs: Sequence[KeypointPair]) -> np.ndarray:
        if len(keypointPairs) < 8:
            raise InvalidOperationException("At least 8 keypoint pairs must be provided")
        xy1, xy2 = _coords(keypointPairs)
        return self.estimate_fundamental_matrices(xy1[None], xy2[None])[0][0]

    # -- :26-94 -----------------------------------------------------------------------------------
    def GetFundamentalMatrix(self, keypointPairs: Sequence[KeypointPair], numSamples: int, numPairsPerSample: int,
                             threshold: float) -> Tuple[List[KeypointPair], np.ndarray]:
        if numPairsPerSample < 8:
            raise InvalidOperationException("At least 8 keypoint pairs must be included per sample")
        if len(keypointPairs) < numPairsPerSample:
            raise InvalidOperationException("Must provide at least as many keypoint pairs as pairs required per sample")
        xy1, xy2 = _coords(keypointPairs)
        rng = np.random.default_rng(self._seed)
        keys = rng.random((numSamples, len(keypointPairs)), dtype=np.float32)        # OrderBy(_ => random.NextSingle())
        idx = np.argsort(keys, axis=1, kind="stable")[:, :numPairsPerSample]         # .Take(numPairsPerSample)
        F, rank = self.estimate_fundamental_matrices(xy1[idx], xy2[idx])
        valid = (rank == 2).astype(np.uint8) if self._require_rank2 else np.ones(numSamples, dtype=np.uint8)
        counts, best, mask = self._matcher.ransac_score(F, xy1, xy2, threshold, valid)
        self.last_hypotheses = (F, valid, counts)
        if best < 0:
            raise Exception("Failed computing the best fundamental matrix")          # :87-88
        return [p for p, keep in zip(keypointPairs, mask.tolist()) if keep], F[best]
