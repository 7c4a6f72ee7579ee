"""numpy restatement of the reference's keypoint producer -- TEST INFRASTRUCTURE ONLY.

The steps immediately upstream of the matcher:

  Grayscale.FromRgba64            dotnet_src/Images.Abstractions/Pixels/Grayscale.cs:19-23
  KeypointDetection (FAST-12)     dotnet_src/ImageProcessing/KeypointDetection.cs:15-138
  Keypoint.GetBriefDescriptor     dotnet_src/ImageProcessing.Abstractions/Keypoint.cs:29-57
  RedundantKeypointEliminator     dotnet_src/ImageProcessing/RedundantKeypointEliminator.cs:16-39
  Utils.NextGaussianCoordinate    dotnet_src/ImageProcessing/Utils.cs:19-38  (structure only: the
                                  reference draws from an UNSEEDED System.Random, so its pair table
                                  cannot be reproduced; here the uniform stream is a seeded splitmix64)

Pinned by the reference's own xUnit tests for FAST (ImageProcessing.Tests/KeypointDetectionTests.cs:10-50),
restated in tests/test_oracle_detect.py.  Images are ``float32[H, W]`` arrays of ``Grayscale.K`` indexed
``img[y, x]`` (the reference indexes ``image[x, y]``).
"""
from __future__ import annotations

import math

import numpy as np

# KeypointDetection.cs:15-19 -- (dx, dy) added to (x, y); the last entry repeats {-3, 1} upstream
# (a typo for {-3, -1}) and is reproduced as is.
BRESENHAM_CIRCLE_3 = np.array([
    (-3, 0), (-3, 1), (-2, 2), (-1, 3), (0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1),
    (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, 1)], dtype=np.int32)
MINI_BRESENHAM_CIRCLE_3 = np.array([(-3, 0), (0, 3), (3, 0), (0, -3)], dtype=np.int32)   # :21-22

# The Python generation's table (python_src/photogrammetry/image_processing/keypoint_detection.py:12-29) is
# (height_offset, width_offset) with the correct last entry; as (dx, dy):
PY_BRESENHAM_CIRCLE_3 = np.array([
    (0, -3), (1, -3), (2, -2), (3, -1), (3, 0), (3, 1), (2, 2), (1, 3), (0, 3), (-1, 3),
    (-2, 2), (-3, 1), (-3, 0), (-3, -1), (-2, -2), (-1, -3)], dtype=np.int32)
PY_MINI_BRESENHAM_CIRCLE_3 = PY_BRESENHAM_CIRCLE_3[[0, 4, 8, 12]]                         # :30


def grayscale_from_rgb8(rgb: np.ndarray) -> np.ndarray:
    """8-bit RGB(A) image -> Grayscale.K as the C# pipeline sees it: ImageSharp widens to Rgba64
    (v * 257) and Grayscale.FromRgba64 computes ((float)R + B + G) / (3 * ushort.MaxValue) in float32."""
    c = rgb[..., :3].astype(np.uint32) * 257
    s = (c[..., 0].astype(np.float32) + c[..., 2].astype(np.float32)) + c[..., 1].astype(np.float32)
    return (s / np.float32(3 * 65535)).astype(np.float32)


def in_threshold(intensity, test, threshold):
    """KeypointDetection.cs:135-138 (float32 arithmetic)."""
    i, t, th = np.float32(intensity), np.float32(test), np.float32(threshold)
    return bool(t > np.float32(i - th) and t < np.float32(i + th))


def is_potential_keypoint(img, intensity, x, y, threshold, python_generation=False):
    """KeypointDetection.cs:116-133 (= _is_keypoint_quick, keypoint_detection.py:72-91): at most one of the
    four compass points may be inside the threshold."""
    inside = 0
    for dx, dy in (PY_MINI_BRESENHAM_CIRCLE_3 if python_generation else MINI_BRESENHAM_CIRCLE_3):
        if not in_threshold(intensity, img[y + dy, x + dx], threshold):
            continue
        if inside > 0:
            return False
        inside += 1
    return True


def intensity_value_if_keypoint(img, x, y, threshold, python_generation=False):
    """KeypointDetection.cs:65-114: longest circular run of ring pixels OUTSIDE the threshold, None if < 12
    or if a fifth inside-threshold pixel is met.  The Python generation's _is_keypoint
    (keypoint_detection.py:93-115) accepts exactly the same pixels on its own ring table: it returns True at
    the first run of 12 (at most 4 ring pixels are then inside, so the fifth-failure exit cannot fire first
    on an accepted pixel) and otherwise tests the wrapped final run."""
    intensity = img[y, x]
    if not is_potential_keypoint(img, intensity, x, y, threshold, python_generation):
        return None
    beginning, n_begin, longest, current, n_fail = True, 0, 0, 0, 0
    for dx, dy in (PY_BRESENHAM_CIRCLE_3 if python_generation else BRESENHAM_CIRCLE_3):
        if in_threshold(intensity, img[y + dy, x + dx], threshold):
            beginning = False
            longest = max(longest, current)
            current = 0
            if n_fail >= 4:
                return None
            n_fail += 1
        else:
            current += 1
            if beginning:
                n_begin += 1
    if not beginning:
        current += n_begin
    longest = max(longest, current)
    return None if longest < 12 else longest


def py_is_keypoint(img, x, y, threshold):
    """Literal restatement of the Python generation's pixel test (_is_keypoint_quick + _is_keypoint,
    keypoint_detection.py:72-115), kept separate from the C# loop to check the equivalence claimed above."""
    lo, hi = img[y, x] - threshold, img[y, x] + threshold
    quick = 0
    for dx, dy in PY_MINI_BRESENHAM_CIRCLE_3:
        t = img[y + dy, x + dx]
        if not (t > lo and t < hi):
            continue
        if quick > 0:
            return False
        quick += 1
    beginning, n_begin, n_consec, n_fail = True, 0, 0, 0
    for dx, dy in PY_BRESENHAM_CIRCLE_3:
        t = img[y + dy, x + dx]
        if t > lo and t < hi:
            beginning = False
            n_consec = 0
            n_fail += 1
            if n_fail > 4:
                return False
        else:
            n_consec += 1
            if beginning:
                n_begin += 1
            if n_consec >= 12:
                return True
    return n_consec + n_begin >= 12


def detect(img: np.ndarray, threshold: float, python_generation: bool = False):
    """KeypointDetection.Detect (:42-63): row-major scan of the interior; returns
    (coords int32[n, 2] as (x, y), scores int32[n])."""
    h, w = img.shape
    coords, scores = [], []
    for y in range(3, h - 3):
        for x in range(3, w - 3):
            s = intensity_value_if_keypoint(img, x, y, threshold, python_generation)
            if s is not None:
                coords.append((x, y))
                scores.append(s)
    return np.array(coords, dtype=np.int32).reshape(-1, 2), np.array(scores, dtype=np.int32)


def detect_vectorised(img: np.ndarray, threshold: float, python_generation: bool = False):
    """Same result as :func:`detect`, vectorised over pixels (for images too large for the scalar loop);
    cross-checked against it in the tests."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    h, w = img.shape
    th = np.float32(threshold)
    c = img[3:h - 3, 3:w - 3]
    lo, hi = (c - th).astype(np.float32), (c + th).astype(np.float32)

    def inside(dx, dy):
        t = img[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx]
        return (t > lo) & (t < hi)

    # IsPotentialKeypoint: fails at the second inside point
    cnt = np.zeros(c.shape, dtype=np.int32)
    for dx, dy in (PY_MINI_BRESENHAM_CIRCLE_3 if python_generation else MINI_BRESENHAM_CIRCLE_3):
        cnt += inside(dx, dy)
    potential = cnt <= 1
    beginning = np.ones(c.shape, dtype=bool)
    n_begin = np.zeros(c.shape, dtype=np.int32)
    longest = np.zeros(c.shape, dtype=np.int32)
    current = np.zeros(c.shape, dtype=np.int32)
    n_fail = np.zeros(c.shape, dtype=np.int32)
    dead = np.zeros(c.shape, dtype=bool)
    for dx, dy in (PY_BRESENHAM_CIRCLE_3 if python_generation else BRESENHAM_CIRCLE_3):
        ins = inside(dx, dy)
        longest = np.where(ins, np.maximum(longest, current), longest)
        dead |= ins & (n_fail >= 4)
        n_fail += ins
        n_begin += (~ins) & beginning
        current = np.where(ins, 0, current + 1)
        beginning &= ~ins
    current = np.where(beginning, current, current + n_begin)
    longest = np.maximum(longest, current)
    ok = potential & ~dead & (longest >= 12)
    ys, xs = np.nonzero(ok)          # row-major: y outer, x inner, like the reference's loops
    return np.stack([xs + 3, ys + 3], axis=1).astype(np.int32), longest[ys, xs].astype(np.int32)


def brief_descriptor(img: np.ndarray, x: int, y: int, pairs: np.ndarray, lsb_first: bool = False) -> int:
    """Keypoint.GetBriefDescriptor (Keypoint.cs:29-57).  pairs: int32[n, 2, 2] = ((dx1, dy1), (dx2, dy2)).
    First pair = most significant bit; a pair with either sample outside the image contributes 0.
    lsb_first: the Python generation's bit order (models/keypoint.py:37-49, pair idx -> 2**idx)."""
    h, w = img.shape
    d = 0
    for idx, ((dx1, dy1), (dx2, dy2)) in enumerate(pairs):
        if not lsb_first:
            d <<= 1
        x1, y1 = x + dx1, y + dy1
        if not (0 <= x1 < w and 0 <= y1 < h):
            continue
        x2, y2 = x + dx2, y + dy2
        if not (0 <= x2 < w and 0 <= y2 < h):
            continue
        if img[y1, x1] < img[y2, x2]:
            d += (1 << idx) if lsb_first else 1
    return d


def brief_descriptors(img: np.ndarray, coords: np.ndarray, pairs: np.ndarray, lsb_first: bool = False):
    return [brief_descriptor(img, int(x), int(y), pairs, lsb_first) for x, y in coords]


def py_pairs_to_xy(pairs_uv: np.ndarray) -> np.ndarray:
    """models/keypoint.py pair table ((du1, dv1), (du2, dv2)) = (row, column) offsets -> ((dx1, dy1), (dx2, dy2))."""
    return np.ascontiguousarray(np.asarray(pairs_uv).reshape(-1, 2, 2)[:, :, ::-1]).astype(np.int32)


def eliminate_redundant(coords: np.ndarray, scores: np.ndarray, radius: int) -> np.ndarray:
    """RedundantKeypointEliminator.EliminateRedundantKeypoints (:16-35): stable sort by FastScore descending,
    then repeatedly keep the head and drop everything within `radius` of it (distance <= radius is dropped).
    Returns the indices (into the input) of the kept keypoints, in the reference's output order."""
    order = np.argsort(-scores.astype(np.int64), kind="stable")     # LINQ OrderByDescending is stable
    alive = list(order)
    kept = []
    while alive:
        head = alive.pop(0)
        kept.append(head)
        hx, hy = coords[head]
        alive = [k for k in alive
                 if math.sqrt(float(coords[k][0] - hx) ** 2 + float(coords[k][1] - hy) ** 2) > radius]
    return np.array(kept, dtype=np.int32)


def eliminate_redundant_vectorised(coords: np.ndarray, scores: np.ndarray, radius: int) -> np.ndarray:
    """Same answer as eliminate_redundant (checked in tests/test_oracle_detect.py), one numpy pass per kept
    keypoint instead of a Python list rebuild: usable at tens of thousands of keypoints."""
    order = np.argsort(-scores.astype(np.int64), kind="stable")
    c = coords.astype(np.float64)
    alive = np.ones(len(order), dtype=bool)
    kept = []
    for head in order:
        if not alive[head]:
            continue
        kept.append(head)
        d = np.sqrt((c[:, 0] - c[head, 0]) ** 2 + (c[:, 1] - c[head, 1]) ** 2)
        alive &= d > radius
    return np.array(kept, dtype=np.int32)


# ---- seeded stand-in for Utils.NextGaussianPair ---------------------------------------------
_GAMMA, _M1, _M2, _MASK = 0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB, (1 << 64) - 1


def _sm64(seed: int, idx: int) -> int:
    z = (seed + (idx + 1) * _GAMMA) & _MASK
    z = ((z ^ (z >> 30)) * _M1) & _MASK
    z = ((z ^ (z >> 27)) * _M2) & _MASK
    return z ^ (z >> 31)


def gaussian_pairs(seed: int, num_pairs: int = 256, stdev: int = 50) -> np.ndarray:
    """Utils.NextGaussianPair (Utils.cs:14-38) with a seeded uniform stream: Marsaglia polar method on
    y1, y2 in [0, 1) (as upstream -- NextDouble is non-negative, so offsets are non-negative), truncated to int."""
    out = np.zeros((num_pairs, 2, 2), dtype=np.int32)
    k = 0
    for p in range(num_pairs):
        for e in range(2):
            while True:
                y1 = (_sm64(seed, k) >> 11) * (1.0 / 9007199254740992.0)
                y2 = (_sm64(seed, k + 1) >> 11) * (1.0 / 9007199254740992.0)
                k += 2
                r2 = y1 * y1 + y2 * y2
                if 0.0 < r2 < 1.0:
                    break
            s = math.sqrt(-2.0 * math.log(r2) / r2)
            out[p, e] = (int(s * y1 * stdev), int(s * y2 * stdev))
    return out
