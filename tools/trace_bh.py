"""Timeline of the batch-hard step inside one CUDA-graph replay (developer aid): globaltimer stamps written by the
MMA thread and one epilogue warp of a few GEMM CTAs (csrc/tc_engine.cuh, trace_stamp) and -- with a library built
with -DEN_FIN_TRACE -- by the two finalize kernels (csrc/batch_losses.cu, fin_stamp / fin_sample):

    python -c "from embeddingnet_b200 import build as b; b.build_variant('build/lib_trace.so', ['-DEN_FIN_TRACE'])"
    EMBEDDINGNET_B200_LIB=build/lib_trace.so python tools/trace_bh.py
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
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
buf = torch.zeros(148 * 64 + 8 + 2048 + 256, dtype=torch.int64, device=dev)
buf[148 * 64 + 0] = buf[148 * 64 + 2] = (1 << 62)     # first-entry slots take the minimum
lib.en_debug_set_bh_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.en_debug_set_bh_trace(ctypes.c_void_p(buf.data_ptr()), 148)
torch.cuda.synchronize()
# the step as ONE CUDA graph (what bench.py times): kernel-to-kernel gaps as in production
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    st.step(emb, labels)
for _ in range(2):
    g.replay()
torch.cuda.synchronize()
buf[:148 * 64].zero_()
buf[148 * 64:].zero_()
buf[148 * 64 + 0] = buf[148 * 64 + 2] = (1 << 62)
flush.zero_()
torch.cuda.synchronize()
g.replay()
torch.cuda.synchronize()
lib.en_debug_set_bh_trace(ctypes.c_void_p(0), 148)
fin = buf[148 * 64:].cpu().numpy()
buf = buf[:148 * 64]
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
print("finalize (same clock, us after the GEMM's first entry): fast kernel first warp in %.1f, last warp out %.1f | "
      "slow kernel first block in %.1f, listed anchors done %.1f, mean written %.1f  (%d anchors on the list)" % (
          (fin[0] - t0) / 1e3, (fin[1] - t0) / 1e3, (fin[2] - t0) / 1e3, (fin[3] - t0) / 1e3 if fin[3] else float("nan"),
          (fin[4] - t0) / 1e3, fin[5]))

smp = fin[8 + 2048:8 + 2048 + 256].reshape(64, 4).astype(np.float64)
smp = smp[(smp > 0).all(axis=1)]
if len(smp):
    rel = (smp - t0) / 1e3
    print("fast kernel, 64 sampled warps (us after the GEMM's first entry; median / max): in %.1f / %.1f | records reduced "
          "%.1f / %.1f | exact distances %.1f / %.1f | out %.1f / %.1f   (warps that went to the list skip the middle)" % (
              np.median(rel[:, 0]), rel[:, 0].max(), np.median(rel[:, 1]), rel[:, 1].max(), np.median(rel[:, 2]),
              rel[:, 2].max(), np.median(rel[:, 3]), rel[:, 3].max()))
n = int(fin[5])
per = fin[8:8 + 2 * min(n, 1024)].reshape(-1, 2)
if len(per):
    us = per[:, 0] / 1e3
    start = (per[:, 1] - fin[2]) / 1e3
    print("slow kernel, per listed anchor: resolve time us min %.1f median %.1f p90 %.1f max %.1f | resolve START after "
          "the kernel's first block entered: min %.1f median %.1f p90 %.1f max %.1f" % (
              us.min(), np.median(us), np.percentile(us, 90), us.max(), start.min(), np.median(start),
              np.percentile(start, 90), start.max()))
    order = np.argsort(start)
    print("  start by list position (every 20th):", [round(float(x), 1) for x in start[::20]])
