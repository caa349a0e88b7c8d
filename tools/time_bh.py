"""Stage times of the headline step (B = 4096, d = 512) from CUDA events recorded inside the library between its
kernels (eager launches, L2 flushed before each step): operand split | distance GEMM | fast finalize | slow finalize."""
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
if len(sys.argv) > 1 and sys.argv[1] == "gauss":   # few near-ties: (almost) no anchor takes the slow path
    torch.manual_seed(0)
    emb = lac.l2_normalize(torch.randn(4096, 512, device=dev)).detach()
    labels = torch.arange(4096, dtype=torch.int32, device=dev)   # no positives at all
st = BatchHardStep(4096, 512, 0.5)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = _lib.load()
for _ in range(5):
    st.step(emb, labels)
lib.en_prof_enable(1)
acc = [0.0] * 4
n = 20
for _ in range(n):
    flush.zero_()
    st.step(emb, labels)
    ms = (ctypes.c_float * 8)()
    cnt = ctypes.c_int(0)
    _lib.check(lib.en_prof_marks_ms(ms, 8, ctypes.byref(cnt)), "en_prof_marks_ms")
    for i in range(cnt.value):
        acc[i] += ms[i]
lib.en_prof_enable(0)
names = ["split", "gemm", "finalize_fast", "finalize_slow"]
print("  ".join("%s %.1f us" % (nm, 1e3 * a / n) for nm, a in zip(names, acc)), " total %.1f us" % (1e3 * sum(acc) / n))
