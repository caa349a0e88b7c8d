"""Median time of the batch-hard step as one CUDA-graph replay (L2 flushed before each), for A/B runs (tools/ab_bh.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from embeddingnet_b200 import losses_and_accuracies as lac, synth  # noqa: E402
from embeddingnet_b200.fused import BatchHardStep  # noqa: E402

dev = torch.device("cuda", 0)
raw, labels = synth.make_device(4096, 512, n_classes=512, rows_per_class=8, noise=0.5, relu=True, device=dev)
emb = lac.l2_normalize(raw).detach()
st = BatchHardStep(4096, 512, 0.5)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    st.step(emb, labels)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    st.step(emb, labels)
ts = []
for i in range(420):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    if i >= 20:
        ts.append(a.elapsed_time(b) * 1e3)
# event time stamps have ~2 us granularity on this part: the MEAN over many steps resolves smaller differences
print("%.2f" % float(np.mean(ts)), "%.2f" % float(np.median(ts)), "%.6f" % st.loss.item())
