import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from embeddingnet_b200 import synth, losses_and_accuracies as lac
from oracle import np_oracle as O

def unit(x):
    ss = np.sum(x.astype(np.float64)**2, axis=1, keepdims=True)
    return (x/np.sqrt(np.maximum(ss,1e-12))).astype(np.float32)
def rel(a,b): return float(np.linalg.norm(a.astype(np.float64)-b.astype(np.float64))/np.linalg.norm(b.astype(np.float64)))

for (ncls, per, d) in [(32,8,128),(16,8,256),(37,9,100)]:
    x, lab = synth.make_numpy(ncls*per, d, n_classes=ncls, rows_per_class=per, noise=0.5, relu=True)
    x = unit(x); lab = lab.astype(np.int64)
    for squared in (False, True):
        ga = O.batch_all_grad_analytic(lab, x, 0.5, squared)
        res = {}
        for name, mp in (("tc", per-1), ("v0", 9 if per-1 <= 8 else per-1)):
            e = torch.tensor(x, device="cuda", requires_grad=True)
            loss = lac.batch_all_triplet_loss(0.5, squared=squared, max_positives=mp)(lab, e)
            loss.backward()
            res[name] = e.grad.cpu().numpy()
        print(ncls, per, d, "squared" if squared else "sqrt", "tc vs oracle %.3e  v0 vs oracle %.3e  tc vs v0 %.3e" % (rel(res["tc"], ga), rel(res["v0"], ga), rel(res["tc"], res["v0"])))
        diff = np.abs(res["tc"].astype(np.float64) - ga)
        rows = np.linalg.norm(diff, axis=1) / (np.linalg.norm(ga, axis=1) + 1e-30)
        print("   worst rows:", np.argsort(-rows)[:5], np.sort(rows)[-5:])
