"""Small-batch bank scans: CUDA-core streaming kernel vs the tensor-core scan, per queries-per-call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from embeddingnet_b200 import synth
from embeddingnet_b200.models import BankKNNClassifier

dev = torch.device("cuda", 0)
n, d = 4_000_000, 512
bank, _ = synth.make_device(n, d, n_classes=100000, noise=0.5, device=dev)
ids = (torch.arange(n, device=dev) % 100000).to(torch.int32)
clf = BankKNNClassifier(5, device=dev).fit_shard(bank, ids, 0, n, classes=np.arange(100000))
q, _ = synth.make_device(1024, d, seed_noise=synth.SEED_QUERY, n_classes=100000, noise=0.5, device=dev)


def t(qs, reps=5):
    for _ in range(2):
        clf.kneighbors_device(qs)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        clf.kneighbors_device(qs)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


print("bank %d x %d fp32 = %.2f GB; HBM floor at 6.5 TB/s: %.2f ms" % (n, d, n * d * 4 / 1e9, n * d * 4 / 6.5e9))
for Q in (1, 2, 4, 8, 16, 32, 64, 128, 256, 1024):
    qs = q[:Q].contiguous()
    line = "Q=%4d" % Q
    if Q <= 8:
        clf.stream_max_q = 8
        line += "  stream %.3f ms" % t(qs)
    clf.stream_max_q = 0
    ms = t(qs)
    line += "  tensor %.3f ms (%.0f q/s, uncertified %d)" % (ms, Q / ms * 1e3, clf.last_uncertified)
    print(line, flush=True)
