"""Loads tests/golden/*.npz back into a ReplanBatch."""
import glob
import os

import numpy as np

from neptune_b200 import config
from neptune_b200.batch import ReplanBatch

HERE = os.path.dirname(os.path.abspath(__file__))
BATCH_KEYS = ("agent_id", "n_int", "coeff_init", "hull_ptr", "hull_xy", "nih0", "st_ptr", "st_xy", "esv_cnt",
              "esv_alpha", "esv_active", "bp_cnt", "bp_xy")


def golden_files():
    return sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def load(path):
    z = np.load(path)
    cfg = os.path.basename(path).split("_")[0]
    par = config(cfg)
    b = ReplanBatch(par=par, n_hull_slots=int(z["n_hull_slots"]), **{k: np.ascontiguousarray(z[k]) for k in BATCH_KEYS})
    b.validate()
    return par, b, z
