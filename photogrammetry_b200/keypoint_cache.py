"""Reader/writer for the Python generation's keypoint cache files (host glue,
SURVEY.md section 8 row f4).

``KeypointCache`` pickles a plain ``list`` of keypoint objects per image
(python_src/photogrammetry/storage/keypoint_cache.py:35-49).  The two frozen
fixtures under data/feature_matching_test/*.dat were written by an older
class layout whose instances carry ``{coord, moment, descriptor}``; the class
itself (``photogrammetry.image_processing.keypoint_detection.KeyPoint``) is
not needed to read them, so the unpickler maps every ``photogrammetry.*``
class to an attribute bag instead of importing the reference package.
"""
from __future__ import annotations

import pickle
import warnings
from typing import List

from .keypoint import Coordinate, Keypoint


class _Bag:
    pass


class _CacheUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] == "photogrammetry":
            return type(name, (_Bag,), {})
        if module.split(".")[0] not in ("numpy", "builtins", "collections", "copyreg", "_codecs"):
            raise pickle.UnpicklingError(f"refusing to load {module}.{name} from a keypoint cache file")
        return super().find_class(module, name)


def load_keypoint_dat(path: str) -> List[Keypoint]:
    """Load one ``<uid>.dat`` cache file into a list of :class:`Keypoint`."""
    with open(path, "rb") as f, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        raw = _CacheUnpickler(f).load()
    out = []
    for kp in raw:
        d = kp.__dict__
        coord = d.get("coord", d.get("_coord"))
        desc = d.get("descriptor", d.get("_descriptor"))
        if desc is None:
            raise ValueError("cache entry has no materialised descriptor")
        out.append(Keypoint(Coordinate(int(coord[0]), int(coord[1])), int(desc),
                            Value=float(d["moment"]) if "moment" in d else None))
    return out


class CachedKeyPoint:
    """Minimal picklable stand-in with the cache's attribute layout."""

    def __init__(self, coord, descriptor, moment=0.0):
        self.coord = list(coord)
        self.moment = moment
        self.descriptor = int(descriptor)


def save_keypoint_dat(path: str, keypoints: List[Keypoint]) -> None:
    """Write a list of keypoints in the cache's pickle layout (protocol 4)."""
    raw = [CachedKeyPoint(k.coord, k.BriefDescriptor, k.Value if isinstance(k.Value, float) else 0.0)
           for k in keypoints]
    with open(path, "wb") as f:
        pickle.dump(raw, f, protocol=4)
