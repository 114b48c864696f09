"""Writes tests/golden/loader_obj.npz: the output of the REFERENCE's WavefrontLoader (oracle/_ref/libifrit_ref_loader.so,
compiled from /root/reference by oracle/Makefile) for the OBJ file tests/test_loaders.py::write_obj(seed=5) generates."""
import os
import sys
import tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_loaders as T  # noqa: E402

with tempfile.TemporaryDirectory() as d:
    r = T._reference(T.write_obj(os.path.join(d, "a.obj")))
np.savez_compressed(T.GOLDEN, **r)
print("wrote", T.GOLDEN, {k: v.shape for k, v in r.items()})
