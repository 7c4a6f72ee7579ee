"""Host mirror of the reference's keypoint producer over the C ABI.

C# generation (same names, argument meaning and output order):

  KeypointDetectionOptions              dotnet_src/ImageProcessing/Options/KeypointDetectionOptions.cs:3-12
  KeypointDetection.Detect              dotnet_src/ImageProcessing/KeypointDetection.cs:42-63
  Keypoint.GetBriefDescriptor           dotnet_src/ImageProcessing.Abstractions/Keypoint.cs:29-57
  RedundantKeypointEliminator           dotnet_src/ImageProcessing/RedundantKeypointEliminator.cs:7-39
  Utils.NextGaussianPair                dotnet_src/ImageProcessing/Utils.cs:14-38
  Grayscale.FromRgba64                  dotnet_src/Images.Abstractions/Pixels/Grayscale.cs:19-23

Python generation:

  FASTKeypointDetector.detect_points    python_src/photogrammetry/image_processing/keypoint_detection.py:36-175
  KeyPoint.descriptor                   python_src/photogrammetry/models/keypoint.py:5-50
  generate_gaussian_pairs               python_src/photogrammetry/models/keypoint.py:52-57

The segment test, the descriptor bits and the suppression run in CUDA kernels
(csrc/pgm_detect.cuh) behind ``pgm_fast_detect`` / ``pgm_brief_describe`` /
``pgm_nms``; nothing here computes on the CPU and nothing imports ``oracle/``.

The reference draws its BRIEF sampling pairs from an unseeded RNG (Utils.cs:11;
keypoint.py:56), so descriptors are only reproducible when the pair table is an
explicit input: both detectors below accept one, and otherwise draw it from a
seeded stream.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from .descriptors import unpack_descriptors
from .keypoint import Coordinate, Keypoint
from .keypoint_matching import Matcher, default_matcher
from .synthetic import _sm64


# ---------------------------------------------------------------------------
# C# generation
# ---------------------------------------------------------------------------
@dataclass
class KeypointDetectionOptions:
    """appsettings.json:19-24 defaults."""
    Threshold: float = 0.1
    GaussianStandardDeviation: int = 50
    NumGaussianPairs: int = 256


@dataclass
class RedundantKeypointEliminationOptions:
    SuppressionRadius: int = 50


class Utils:
    """Utils.cs:7-38 with a SEEDED uniform stream in place of ``new Random()``."""

    def __init__(self, seed: int = 0):
        self._seed = int(seed)
        self._k = 0

    def _next_double(self) -> float:
        v = int(_sm64(self._seed, np.uint64(self._k))) >> 11
        self._k += 1
        return v * (1.0 / 9007199254740992.0)

    def NextGaussianCoordinate(self, standardDeviation: int) -> Coordinate:
        # Marsaglia polar method on [0, 1)^2 exactly as upstream (:28-37): offsets come out non-negative
        while True:
            y1, y2 = self._next_double(), self._next_double()
            r2 = y1 * y1 + y2 * y2
            if 0.0 < r2 < 1.0:
                break
        s = math.sqrt(-2.0 * math.log(r2) / r2)
        return Coordinate(int(s * y1 * standardDeviation), int(s * y2 * standardDeviation))

    def NextGaussianPair(self, standardDeviation: int):
        return self.NextGaussianCoordinate(standardDeviation), self.NextGaussianCoordinate(standardDeviation)


def grayscale_from_rgb8(rgb: np.ndarray) -> np.ndarray:
    """8-bit RGB(A) -> ``Grayscale.K`` float32[H, W]: ImageSharp widens a byte to Rgba64 by v*257 and
    Grayscale.FromRgba64 (:19-23) evaluates ((float)R + B + G) / (3 * 65535) in single precision."""
    c = np.asarray(rgb)[..., :3].astype(np.uint32) * 257
    s = (c[..., 0].astype(np.float32) + c[..., 2].astype(np.float32)) + c[..., 1].astype(np.float32)
    return (s / np.float32(3 * 65535)).astype(np.float32)


def _pairs_to_table(pairs: Sequence) -> np.ndarray:
    """[(Coordinate, Coordinate)] or int[n, 2, 2] of (X, Y) -> int32[n, 4] = (dx1, dy1, dx2, dy2)."""
    if isinstance(pairs, np.ndarray):
        return np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 4)
    return np.array([[a.X, a.Y, b.X, b.Y] for a, b in pairs], dtype=np.int32).reshape(-1, 4)


class KeypointDetection:
    """``ImageProcessing.KeypointDetection`` (KeypointDetection.cs:10-139).

    ``Detect(image)`` takes ``Grayscale.K`` as ``float32[H, W]`` (``image[y, x]``; the C# matrix is indexed
    ``[x, y]``) and returns the reference's keypoint list: row-major order, ``FastScore`` = longest run,
    ``BriefDescriptor`` from the pair table, ``Value`` = the centre intensity."""

    def __init__(self, options: Optional[KeypointDetectionOptions] = None, gaussian_pairs=None,
                 seed: int = 0, device: int = 0, matcher: Optional[Matcher] = None):
        self._options = options or KeypointDetectionOptions()
        if gaussian_pairs is None:
            utils = Utils(seed)
            gaussian_pairs = [utils.NextGaussianPair(self._options.GaussianStandardDeviation)
                              for _ in range(self._options.NumGaussianPairs)]
        self._gaussianKeypairs = gaussian_pairs
        self._pair_table = _pairs_to_table(gaussian_pairs)
        self._matcher = matcher or default_matcher(device)

    @property
    def gaussian_pairs(self):
        return self._gaussianKeypairs

    def detect_arrays(self, image: np.ndarray):
        """(xy int32[n, 2], score int32[n], desc uint8[n, stride]) without building objects."""
        image = np.ascontiguousarray(image, dtype=np.float32)
        xy, score = self._matcher.fast_detect(image, self._options.Threshold)
        desc = self._matcher.brief_describe(image, xy, self._pair_table)
        return xy, score, desc

    def Detect(self, image: np.ndarray) -> List[Keypoint]:
        image = np.ascontiguousarray(image, dtype=np.float32)
        xy, score, desc = self.detect_arrays(image)
        ints = unpack_descriptors(desc)
        return [Keypoint(Coordinate=Coordinate(int(x), int(y)), BriefDescriptor=d, FastScore=int(s),
                         Value=float(image[y, x]))
                for (x, y), s, d in zip(xy.tolist(), score.tolist(), ints)]


class RedundantKeypointEliminator:
    """``ImageProcessing.RedundantKeypointEliminator`` (RedundantKeypointEliminator.cs:7-39)."""

    def __init__(self, options: Optional[RedundantKeypointEliminationOptions] = None, device: int = 0,
                 matcher: Optional[Matcher] = None):
        self._suppressionRadius = (options or RedundantKeypointEliminationOptions()).SuppressionRadius
        self._matcher = matcher or default_matcher(device)

    def EliminateRedundantKeypoints(self, keypoints: List[Keypoint]) -> List[Keypoint]:
        if not keypoints:
            return []
        xy = np.array([[k.Coordinate.X, k.Coordinate.Y] for k in keypoints], dtype=np.int32)
        score = np.array([k.FastScore for k in keypoints], dtype=np.int32)
        kept = self._matcher.nms(xy, score, self._suppressionRadius)
        return [keypoints[i] for i in kept.tolist()]          # the caller's own objects, reference order


# ---------------------------------------------------------------------------
# Python generation
# ---------------------------------------------------------------------------
def generate_gaussian_pairs(stdev, num_pairs: int = 256) -> np.ndarray:
    """keypoint.py:52-57: ``int64[num_pairs, 2, 2]`` of (height_offset, width_offset), drawn from numpy's
    GLOBAL generator in the reference's order, so ``np.random.seed(s)`` reproduces the reference's table."""
    table = np.zeros((num_pairs, 2, 2), dtype=np.int64)
    for row in table:
        for end in range(2):
            row[end] = np.rint(np.random.normal([0, 0], stdev)).astype(np.int64)
    return table


class KeyPoint:
    """models/keypoint.py:5-30: ``coord`` = [u, v] = [row, column], ``descriptor`` = int, bit idx = pair idx."""

    def __init__(self, image_id: int, coord: np.ndarray, descriptor: int):
        self._image_id = image_id
        self._coord = coord
        self._descriptor = int(descriptor)

    @property
    def coord(self):
        return self._coord

    @property
    def descriptor(self) -> int:
        return self._descriptor


class FASTKeypointDetector:
    """``FASTKeypointDetector`` (keypoint_detection.py:36-175).  ``image_db`` is anything with
    ``get_bw_image(image_id) -> int16[H, W]`` (storage/image_db.py:32-37), or the gray image itself."""

    def __init__(self, threshold, image_db, gaussian_pairs: Optional[np.ndarray] = None, device: int = 0,
                 matcher: Optional[Matcher] = None):
        self.threshold = threshold
        self._image_db = image_db
        self._gaussian_pairs = generate_gaussian_pairs(stdev=50) if gaussian_pairs is None else np.asarray(gaussian_pairs)
        self._matcher = matcher or default_matcher(device)

    def _bw(self, image_id) -> np.ndarray:
        img = self._image_db.get_bw_image(image_id) if hasattr(self._image_db, "get_bw_image") else self._image_db
        img = np.asarray(img)
        if img.ndim != 2:
            raise ValueError("expected a single-channel image")
        return img.astype(np.float32)      # int16 gray values and an integer threshold are exact in float32

    def detect_arrays(self, image_id: int = 0):
        """(uv int32[n, 2] as (row, column), desc uint8[n, stride])."""
        bw = self._bw(image_id)
        xy, _ = self._matcher.fast_detect(bw, float(self.threshold), python_generation=True)
        gp = self._gaussian_pairs.reshape(-1, 2, 2)
        table = np.stack([gp[:, 0, 1], gp[:, 0, 0], gp[:, 1, 1], gp[:, 1, 0]], axis=1).astype(np.int32)
        desc = self._matcher.brief_describe(bw, xy, table, python_generation=True)
        return xy[:, ::-1].copy(), desc

    def detect_points(self, image_id: int = 0) -> List[KeyPoint]:
        uv, desc = self.detect_arrays(image_id)
        return [KeyPoint(image_id, np.array(c), d) for c, d in zip(uv.tolist(), unpack_descriptors(desc))]
