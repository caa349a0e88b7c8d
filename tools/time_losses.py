"""Times the in-batch losses (fwd, fwd+bwd) at BASELINE config 3 through the public API (CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from embeddingnet_b200 import synth, losses_and_accuracies as lac

dev = torch.device("cuda", 0)
B, d = 4096, 512
raw, labels = synth.make_device(B, d, n_classes=512, rows_per_class=8, noise=0.5, relu=True, device=dev)
emb = lac.l2_normalize(raw).detach()

def timeit(name, fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-40s %9.3f ms" % (name, e0.elapsed_time(e1) / n))

def fwd_bwd(loss_fn, x):
    e = x.clone().requires_grad_(True)
    l = loss_fn(labels, e)
    l.backward()
    return e.grad

bh = lac.batch_hard_triplet_loss(0.5)
ba = lac.batch_all_triplet_loss(0.5, max_positives=7)
ca = lac.contrastive_loss_all_pairs()
x7 = (emb * 0.7).contiguous()
timeit("batch_hard fwd", lambda: bh(labels, emb))
timeit("batch_hard fwd+bwd", lambda: fwd_bwd(bh, emb))
timeit("batch_all fwd", lambda: ba(labels, emb))
timeit("batch_all fwd+bwd", lambda: fwd_bwd(ba, emb))
timeit("contrastive_all_pairs fwd", lambda: ca(labels, x7))
timeit("contrastive_all_pairs fwd+bwd", lambda: fwd_bwd(ca, x7))
y = torch.cat([emb, emb.roll(1, 0), emb.roll(9, 0)], dim=1).contiguous()
tl = lac.triplet_loss(0.5)
timeit("triplet_loss [a|p|n] fwd (25 MB)", lambda: tl(None, y))
timeit("l2_normalize fwd (8 MB in, 8 out)", lambda: lac.l2_normalize(raw))
