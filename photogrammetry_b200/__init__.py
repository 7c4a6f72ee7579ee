"""photogrammetry_b200 -- B200-native descriptor matching for Takatsuka-Mark/Photogrammetry.

One hot path, rebuilt from scratch for sm_100a behind the reference's own
surface: ``ImageProcessing.KeypointMatching.MatchKeypoints``
(dotnet_src/ImageProcessing/KeypointMatching.cs:14-69) -- the brute-force
Hamming distance matrix between two images' BRIEF descriptors followed by the
greedy one-to-one assignment.  The compute lives in ``csrc/`` (CUDA kernels +
the C-ABI of include/pgmatch.h, built into ``libpgmatch.so``); this package is
the thin host mirror of the reference interface over that C-ABI.

There is no CPU fallback: every compute entry point raises
``PgmatchLibraryError`` if ``libpgmatch.so`` is missing or no GPU is present.
"""
from .descriptors import pack_descriptors, stride_for_bits, unpack_descriptors
from .keypoint import Coordinate, Keypoint, KeypointPair

__all__ = [
    "Coordinate", "Keypoint", "KeypointPair",
    "pack_descriptors", "unpack_descriptors", "stride_for_bits",
]
__version__ = "0.1.0"
