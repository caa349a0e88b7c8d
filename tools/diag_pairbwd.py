"""Diagnostic: batch-all / contrastive gradient error statistics at C3 vs the float64 oracle (not a test)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np, torch
from conftest import unit_rows
from embeddingnet_b200 import synth, losses_and_accuracies as lac
from oracle import np_oracle as O

x, lab = synth.make_numpy(4096, 512, n_classes=512, rows_per_class=8, noise=0.5, relu=True)
x = unit_rows(x)
perm = np.random.RandomState(0).permutation(len(lab))
x, lab = x[perm], lab[perm].astype(np.int64)
ga = O.batch_all_grad_analytic(lab, x, 0.5, False)
scale = np.linalg.norm(ga, axis=1).mean()
prev = None
for it in range(3):
    e = torch.tensor(x, device="cuda", requires_grad=True)
    loss = lac.batch_all_triplet_loss(0.5, max_positives=7)(lab, e)
    loss.backward()
    g = e.grad.cpu().numpy().astype(np.float64)
    rows = np.linalg.norm(g - ga, axis=1) / scale
    print("batch_all it%d loss %.9f median %.3e frac<1e-4 %.4f max %.3e total %.3e same_as_prev %s" % (
        it, loss.item(), np.median(rows), np.mean(rows < 1e-4), rows.max(),
        np.linalg.norm(g - ga) / np.linalg.norm(ga), None if prev is None else bool((prev == g).all())))
    prev = g
x7 = (x * 0.7).astype(np.float32)
gc = O.contrastive_allpairs_grad_analytic(lab, x7)
for it in range(2):
    e7 = torch.tensor(x7, device="cuda", requires_grad=True)
    lc = lac.contrastive_loss_all_pairs()(lab, e7)
    lc.backward()
    g = e7.grad.cpu().numpy().astype(np.float64)
    print("contrastive it%d rel %.3e" % (it, np.linalg.norm(g - gc) / np.linalg.norm(gc)))
