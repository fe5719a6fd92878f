import glob
import os

import numpy as np

import dpgo_b200 as D

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = D.PoseGraph(int(z["d"]), int(z["num_poses"]), z["i"], z["j"], z["R"], z["t"], z["kappa"], z["tau"])
    return g, z
