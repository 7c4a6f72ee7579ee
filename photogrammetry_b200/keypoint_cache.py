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


# The only globals a keypoint cache file legitimately names (the frozen fixtures use numpy.dtype and
# numpy.core.multiarray.scalar for the ``moment`` field; arrays and plain containers are allowed for caches written by
# newer layouts).  Everything else -- builtins.eval, os.system, numpy helpers that call back into Python -- is refused:
# whitelisting whole modules would let a crafted file reach them through REDUCE.
_ALLOWED_GLOBALS = frozenset({
    ("numpy", "dtype"), ("numpy", "ndarray"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("builtins", "list"), ("builtins", "dict"), ("builtins", "tuple"), ("builtins", "set"),
    ("builtins", "int"), ("builtins", "float"), ("builtins", "bytes"), ("builtins", "bytearray"),
    ("_codecs", "encode"),            # how protocol <= 2 pickles spell bytes
})


class _CacheUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] == "photogrammetry":
            return type(name, (_Bag,), {})
        if (module, name) not in _ALLOWED_GLOBALS:
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


def save_keypoint_dat(path: str, keypoints: List[Keypoint]) -> None:
    """Write keypoints the way the Python generation's ``KeypointCache`` does
    (keypoint_cache.py:41-49: ``pickle.dump(list_of_KeyPoint)``), so the file can be
    read back by the reference's own ``pickle.load`` with its package importable:
    instances of ``photogrammetry.models.keypoint.KeyPoint`` carrying ``_coord`` and a
    materialised ``_descriptor`` (models/keypoint.py:5-30)."""
    import sys
    import types

    mod_name = "photogrammetry.models.keypoint"
    created = []
    parts = mod_name.split(".")
    for k in range(1, len(parts) + 1):
        name = ".".join(parts[:k])
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
            created.append(name)
    mod = sys.modules[mod_name]
    had = hasattr(mod, "KeyPoint")
    if not had:
        mod.KeyPoint = type("KeyPoint", (), {"__module__": mod_name})
    try:
        raw = []
        for k in keypoints:
            o = mod.KeyPoint.__new__(mod.KeyPoint)
            o.__dict__.update({"_image_id": 0, "_coord": [int(k.coord[0]), int(k.coord[1])],
                               "_descriptor": int(k.BriefDescriptor), "_gaussian_pairs": None, "_image_db": None})
            raw.append(o)
        with open(path, "wb") as f:
            pickle.dump(raw, f, protocol=4)
    finally:
        if not had:
            del mod.KeyPoint
        for name in created:
            sys.modules.pop(name, None)
