"""ms per step of the host-buffer pipeline (en_bh_host_pipe_*) for several depths / lags, three repetitions each
(A/B two builds on one box:  EMBEDDINGNET_B200_LIB=build/lib_x.so python tools/time_pipe.py)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from embeddingnet_b200 import _lib, synth  # noqa: E402
from embeddingnet_b200.fused import BatchHardHostPipeline  # noqa: E402

B, D = 4096, 512
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
_lib.load()
emb, labels = synth.make_device(B, D, n_classes=512, rows_per_class=8, noise=0.5, relu=True, device=dev)
emb = torch.nn.functional.normalize(emb, dim=1).contiguous()
emb_h = emb.cpu().pin_memory()
lab_h = labels.cpu().pin_memory()
print("library:", os.environ.get("EMBEDDINGNET_B200_LIB", "(default)"))
for depth, lag in ((3, 2), (4, 3), (5, 4), (6, 5), (8, 7), (8, 4)):
    pipe = BatchHardHostPipeline(B, D, margin=0.5, depth=depth)
    pg = [BatchHardHostPipeline.pinned((B, D)) for _ in range(depth)]
    pl = [BatchHardHostPipeline.pinned((1,)) for _ in range(depth)]
    base = 0

    def run(n):
        global base
        for i in range(n + lag):
            if i < n:
                pipe.submit(emb_h, lab_h, pl[i % depth], pg[i % depth])
            j = i - lag
            if j >= 0:
                pipe.wait(base + j)
                float(pl[j % depth][0])
        base += n

    run(20)
    out = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run(300)
        torch.cuda.synchronize()
        out.append((time.perf_counter() - t0) / 300 * 1e3)
    print("depth %d, results read %d steps behind: %s ms/step" % (depth, lag, " ".join("%.4f" % x for x in out)))
    pipe.close()
