"""Batch-all loss + gradient at the stress shape of SURVEY 8(d) (64 classes x 64 rows, B = 4096, d = 512): ms per
fwd+bwd through the autograd callable, and the C3 shape (512 x 8) beside it."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from embeddingnet_b200 import losses_and_accuracies as lac, synth  # noqa: E402

dev = torch.device("cuda", 0)
for ncls, per, mp in ((64, 64, 63), (256, 16, 15), (512, 8, 7)):
    raw, lab = synth.make_device(4096, 512, n_classes=ncls, rows_per_class=per, noise=0.5, relu=True, device=dev)
    emb = lac.l2_normalize(raw).detach().contiguous()
    fn = lac.batch_all_triplet_loss(0.5, max_positives=mp, return_fraction=True)

    def step():
        e = emb.detach().clone().requires_grad_(True)
        loss, frac = fn(lab, e)
        loss.backward()
        return loss, frac

    for _ in range(3):
        loss, frac = step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        step()
    b.record()
    torch.cuda.synchronize()
    print("%4d classes x %2d: %.3f ms per fwd+bwd   loss %.6f  positive fraction %.4f" % (
        ncls, per, a.elapsed_time(b) / 10, loss.item(), frac.item()))
