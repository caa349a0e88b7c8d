"""Times the batch-hard step with and without the gradient (same GEMM, finalize with / without the scatter)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from embeddingnet_b200 import _lib, synth, losses_and_accuracies as lac
from embeddingnet_b200._runtime import ptr, stream_ptr
from embeddingnet_b200.fused import BatchHardStep

dev = torch.device("cuda", 0)
B, D = 4096, 512
raw, labels = synth.make_device(B, D, n_classes=512, rows_per_class=8, noise=0.5, relu=True, device=dev)
emb = lac.l2_normalize(raw).detach().contiguous()
st = BatchHardStep(B, D, 0.5)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def fwd_only():
    si, sf = st.saved_i, st.saved_f
    _lib.call("en_batch_hard_fwd", ptr(emb), ptr(labels), B, D, ctypes.c_float(0.5), 0, 0, ptr(st.loss), ptr(si[0]),
              ptr(si[1]), ptr(sf[0]), ptr(sf[1]), ptr(sf[2]), ptr(st.ws), st.ws.numel(), stream_ptr())


def graph_time(fn, n=30):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); s.synchronize()
        with torch.cuda.graph(g, stream=s):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


print("fwd+bwd step: %.1f us" % graph_time(lambda: st.step(emb, labels)))
print("fwd only    : %.1f us" % graph_time(fwd_only))
# shuffled label order: the label-range shortcut of the epilogue never applies
perm = torch.randperm(B, device=dev)
emb2, lab2 = emb[perm].contiguous(), labels[perm].contiguous()
print("fwd+bwd step, shuffled rows: %.1f us" % graph_time(lambda: st.step(emb2, lab2)))
