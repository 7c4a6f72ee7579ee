"""Host side of the configs[2] proxy: the frame construction of python_src/scripts/image_editing.py:8-15 (CPU only)."""
import os

import numpy as np

from photogrammetry_b200 import sequence

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_shifted_frames_follow_the_reference_script():
    g = sequence.star_gray()
    fr = sequence.shifted_frames(g, 32, 5)
    assert fr.shape == (32,) + g.shape and (fr[0] == g).all()
    h, w = g.shape
    off = 150                                     # the offset scripts/image_editing.py uses: frame 30 of the sequence
    exp = np.zeros_like(g)
    for col in range(w - off):                    # image_editing.py:11-13
        exp[:, col + off] = g[:, col]
    assert (fr[30] == exp).all()
    star = np.load(os.path.join(GOLDEN, "star_detect.npz"))
    assert (fr[30] == star["gray1"]).all()        # = 15pt_star_shifted_150.png as shipped by the reference


def test_pairs_table_layout():
    p = sequence.star_pairs()
    t = sequence.pairs_table(p)
    assert t.shape == (256, 4) and t.dtype == np.int32
    assert (t[:, 0] == p[:, 0, 1]).all() and (t[:, 1] == p[:, 0, 0]).all()     # (row, column) offsets -> (dx, dy)
    assert (sequence.pairs_table(t) == t).all()
