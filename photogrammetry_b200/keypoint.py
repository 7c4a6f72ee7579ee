"""Host-side mirrors of the reference's domain types on the matching path.

  Keypoint      dotnet_src/ImageProcessing.Abstractions/Keypoint.cs:9-15
                (Coordinate, FastScore, BriefDescriptor : BigInteger)
  KeypointPair  dotnet_src/ImageProcessing.Abstractions/KeypointPair.cs:3-8
                (Keypoint1, Keypoint2, int Distance)
  Coordinate    dotnet_src/Math/LinearAlgebra/Coordinate.cs:3-17

Member names keep the reference's C# spelling so code written against the
reference reads the same here.  The matcher only ever reads
``BriefDescriptor``; everything else rides along by reference, exactly as the
C# ``KeypointPair`` holds references into the two input lists
(KeypointMatching.cs:57-62).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional


@dataclass(frozen=True)
class Coordinate:
    X: int
    Y: int

    def Add(self, other: "Coordinate") -> "Coordinate":
        return Coordinate(self.X + other.X, self.Y + other.Y)


@dataclass(eq=False)
class Keypoint:
    Coordinate: Coordinate
    BriefDescriptor: int
    FastScore: int = 0
    Value: Optional[Any] = None

    # python_src/photogrammetry/models/keypoint.py spelling, so lists unpickled
    # from the Python generation's KeypointCache can be passed straight in.
    @property
    def descriptor(self) -> int:
        return self.BriefDescriptor

    @property
    def coord(self):
        return (self.Coordinate.X, self.Coordinate.Y)

    def __str__(self) -> str:  # Keypoint.cs:63-66
        return f"({self.Coordinate.X}, {self.Coordinate.Y})"


@dataclass(eq=False)
class KeypointPair:
    Keypoint1: Keypoint
    Keypoint2: Keypoint
    Distance: int


INT_MAX_DISTANCE = 2147483647  # int.MaxValue, KeypointMatching.cs:40
