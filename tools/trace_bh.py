"""Timeline of the batch-hard distance GEMM inside one step (developer aid): globaltimer stamps written by the MMA
thread and one epilogue warp of a few CTAs (csrc/tc_engine.cuh, trace_stamp)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from embeddingnet_b200 import _lib, losses_and_accuracies as lac, synth  # noqa: E402
from embeddingnet_b200.fused import BatchHardStep  # noqa: E402

dev = torch.device("cuda", 0)
raw, labels = synth.make_device(4096, 512, n_classes=512, rows_per_class=8, noise=0.5, relu=True, device=dev)
emb = lac.l2_normalize(raw).detach()
st = BatchHardStep(4096, 512, 0.5)
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    st.step(emb, labels)
buf = torch.zeros(148 * 64, dtype=torch.int64, device=dev)
lib.en_debug_set_bh_trace.argtypes = [ctypes.c_void_p]
lib.en_debug_set_bh_trace(ctypes.c_void_p(buf.data_ptr()))
flush.zero_()
st.step(emb, labels)
torch.cuda.synchronize()
lib.en_debug_set_bh_trace(ctypes.c_void_p(0))
t = buf.cpu().numpy().reshape(148, 64)
t0 = t[:, 0][t[:, 0] > 0].min()
for cta in (0, 1, 73, 147):
    r = t[cta]
    rel = lambda x: (x - t0) / 1e3 if x > 0 else float("nan")  # noqa: E731
    print("CTA %3d: entry %.1f us, set-up done %.1f" % (cta, rel(r[0]), rel(r[1])))
    for k in range(5):
        if r[8 + 4 * k] == 0:
            break
        print("   tile %d: acc free %.1f | operands %.1f | MMAs issued %.1f || epilogue: acc full %.1f, drained %.1f" % (
            k, rel(r[8 + 4 * k]), rel(r[9 + 4 * k]), rel(r[10 + 4 * k]), rel(r[40 + 2 * k]), rel(r[41 + 2 * k])))
last = t[:, 40:60].max()
print("last epilogue stamp over all CTAs: %.1f us after the first entry" % ((last - t0) / 1e3))
