"""Working version of the reference CLI ``tools/test.py`` (broken in the snapshot: SURVEY.md D4) over the new API.

    python tools/test.py config.yml --encodings encodings.pkl --image img.png [--weights model.pt]
    python tools/test.py config.yml --encodings encodings.pkl --encoding query.npy        # skip the backbone

Same arguments as the reference (test.py:5-13).  The backbone is out of scope of this repo: ``--weights`` may point
to a TorchScript module mapping a (1, H, W, 3) float tensor to a (1, d) embedding; with ``--encoding`` a saved query
embedding is classified directly against the bank.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402


class _TorchScriptModel:
    def __init__(self, path):
        import torch

        self.m = torch.jit.load(path).eval().cuda()

    def predict(self, imgs):
        import torch

        with torch.no_grad():
            return self.m(torch.as_tensor(np.asarray(imgs), dtype=torch.float32).cuda()).cpu().numpy()


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("config", type=str, help="path to config file")
    parser.add_argument("--weights", type=str, help="path to trained model weights file (TorchScript)")
    parser.add_argument("--encodings", type=str, help="path to trained model encodings file")
    parser.add_argument("--image", type=str, help="path to image file")
    parser.add_argument("--encoding", type=str, help="path to a .npy query embedding (instead of --image)")
    opt = parser.parse_args()

    import yaml

    from embeddingnet_b200.models import EmbeddingNet

    with open(opt.config) as f:
        cfg = yaml.safe_load(f)
    params = {"model": {k.lower(): v for k, v in cfg.get("MODEL", {}).items()},
              "encodings": {"knn_k": cfg.get("ENCODINGS", {}).get("knn_k", 5)},
              "general": {}}
    model = EmbeddingNet(params, base_model=_TorchScriptModel(opt.weights) if opt.weights else None)
    model.load_encodings(opt.encodings)
    if opt.encoding:
        model_prediction = model.predict_encoding(np.load(opt.encoding))
    else:
        model_prediction = model.predict(opt.image)
    print("Model prediction: {}".format(model_prediction))
