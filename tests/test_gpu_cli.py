"""tools/test.py end to end: the reference CLI (tools/test.py:5-25: config, --weights, --encodings, --image) over a
pickled encoding bank, run as a subprocess exactly as a user would."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from embeddingnet_b200 import synth
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cli_predicts_from_a_pickled_bank(tmp_path, lib_built):
    bank, labels = synth.make_numpy(500, 64, n_classes=25, noise=0.4)
    names = ["sign_%02d" % l for l in labels]
    with open(tmp_path / "encodings.pkl", "wb") as f:      # the reference's layout (models.py:80-90)
        pickle.dump({"paths": ["%d.png" % i for i in range(500)], "labels": names, "encodings": bank}, f)
    (tmp_path / "cfg.yml").write_text("MODEL:\n  input_shape: [48, 48, 3]\nENCODINGS:\n  knn_k: 5\n")
    q, _ = synth.make_numpy(3, 64, seed_noise=synth.SEED_QUERY, n_classes=25, noise=0.4)
    for i in range(3):
        np.save(tmp_path / "q.npy", q[i])
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "test.py"), str(tmp_path / "cfg.yml"),
                            "--encodings", str(tmp_path / "encodings.pkl"), "--encoding", str(tmp_path / "q.npy")],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        assert r.stdout.strip().splitlines()[-1] == "Model prediction: %s" % O.predict_1nn(bank, names, q[i])


def test_cli_rejects_missing_bank(tmp_path, lib_built):
    (tmp_path / "cfg.yml").write_text("MODEL: {}\n")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "test.py"), str(tmp_path / "cfg.yml"),
                        "--encodings", str(tmp_path / "nope.pkl"), "--encoding", str(tmp_path / "q.npy")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "nope.pkl" in r.stderr
