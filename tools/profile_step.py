"""Small driver for ncu: a few batch-hard steps at C3 plus a reduced bank scan (kept short: ncu replays kernels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from embeddingnet_b200 import synth, losses_and_accuracies as lac
from embeddingnet_b200.fused import BatchHardStep
from embeddingnet_b200.models import BankKNNClassifier
import numpy as np

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda", 0)
if what in ("all", "triplet"):
    raw, labels = synth.make_device(4096, 512, n_classes=512, rows_per_class=8, noise=0.5, relu=True, device=dev)
    emb = lac.l2_normalize(raw).detach()
    st = BatchHardStep(4096, 512, 0.5)
    for _ in range(4):
        st.step(emb, labels)
    torch.cuda.synchronize()
if what == "knn_full":
    # one full-size C5 scan (100k queries vs 10M x 512) for the DRAM-traffic metric
    n, Q = 10_000_000, 100_000
    bank, _ = synth.make_device(n, 512, n_classes=100000, noise=0.5, device=dev)
    ids = (torch.arange(n, device=dev) % 100000).to(torch.int32)
    clf = BankKNNClassifier(5, device=dev).fit_shard(bank, ids, 0, n, classes=np.arange(100000))
    q, _ = synth.make_device(Q, 512, seed_noise=synth.SEED_QUERY, n_classes=100000, noise=0.5, device=dev)
    clf.kneighbors_device(q)
    clf.kneighbors_device(q[:8].contiguous())
    clf.kneighbors_device(q[:1].contiguous())
    torch.cuda.synchronize()
if what == "pairbwd":
    raw, labels = synth.make_device(4096, 512, n_classes=512, rows_per_class=8, noise=0.5, relu=True, device=dev)
    emb = lac.l2_normalize(raw).detach()
    ba = lac.batch_all_triplet_loss(0.5, max_positives=7)
    ca = lac.contrastive_loss_all_pairs()
    for fn, x in ((ba, emb), (ca, (emb * 0.7).contiguous())):
        for _ in range(2):
            e = x.clone().requires_grad_(True)
            fn(labels, e).backward()
    torch.cuda.synchronize()
if what == "ba64":
    # stress shape of SURVEY 8(d): 64 classes x 64 rows (63 positives per anchor), batch-all loss + gradient
    raw, labels = synth.make_device(4096, 512, n_classes=64, rows_per_class=64, noise=0.5, relu=True, device=dev)
    emb = lac.l2_normalize(raw).detach()
    ba = lac.batch_all_triplet_loss(0.5, max_positives=63)
    for _ in range(2):
        e = emb.clone().requires_grad_(True)
        ba(labels, e).backward()
    torch.cuda.synchronize()
if what == "mining":
    # C4-shaped bank mining at reduced size: hardest (label-excluded 1-NN) and semihard (count + select passes)
    n, A = 200_000, 16384
    bank, _ = synth.make_device(n, 256, n_classes=2000, noise=0.5, device=dev)
    bank = lac.l2_normalize(bank).detach().contiguous()
    ids = (torch.arange(n, device=dev) % 2000).to(torch.int32)
    clf = BankKNNClassifier(1, device=dev).fit_shard(bank, ids, 0, n, classes=np.arange(2000))
    a_idx = torch.arange(0, n, n // A, device=dev)[:A]
    anchors, a_lab = bank[a_idx].contiguous(), ids[a_idx].contiguous()
    pos = bank[(a_idx + 2000) % n].unsqueeze(1).contiguous()
    np.random.seed(0)
    for mode in ("hardest", "semihard"):
        clf.mine_negatives(anchors, a_lab, positives=pos, margin=0.5, mode=mode)
    torch.cuda.synchronize()
if what == "smallq":
    # the reference's small call patterns against a 4M x 512 bank: 1 query (streaming scan) and 8 / 32 queries
    # (bank-stationary tensor scan)
    n = 4_000_000
    bank, _ = synth.make_device(n, 512, n_classes=100000, noise=0.5, device=dev)
    ids = (torch.arange(n, device=dev) % 100000).to(torch.int32)
    clf = BankKNNClassifier(5, device=dev).fit_shard(bank, ids, 0, n, classes=np.arange(100000))
    q, _ = synth.make_device(64, 512, seed_noise=synth.SEED_QUERY, n_classes=100000, noise=0.5, device=dev)
    for qn in (1, 8, 32):
        for _ in range(2):
            clf.kneighbors_device(q[:qn].contiguous())
    torch.cuda.synchronize()
if what == "rowwise":
    from embeddingnet_b200 import _lib
    from embeddingnet_b200._runtime import ptr, stream_ptr
    rows, d = 1_000_000, 256
    x = synth.make_device(rows, d, n_classes=rows // 8, rows_per_class=8, noise=0.5, relu=True, device=dev)[0]
    g = synth.make_device(rows, d, seed_noise=77, device=dev)[0]
    y = torch.empty_like(x)
    v = torch.empty(rows, device=dev)
    for _ in range(2):
        _lib.call("en_l2_normalize_fwd", ptr(x), ptr(y), rows, d, stream_ptr())
        _lib.call("en_l2_normalize_bwd", ptr(x), ptr(g), ptr(y), rows, d, stream_ptr())
        _lib.call("en_siamese_l2_fwd", ptr(x), ptr(g), rows, d, ptr(v), stream_ptr())
        _lib.call("en_query_distances", ptr(x), ptr(g), rows, d, ptr(v), stream_ptr())
    torch.cuda.synchronize()
if what in ("all", "knn"):
    n, Q = 400_000, 8192
    bank, _ = synth.make_device(n, 512, n_classes=100000, noise=0.5, device=dev)
    ids = (torch.arange(n, device=dev) % 100000).to(torch.int32)
    clf = BankKNNClassifier(5, device=dev).fit_shard(bank, ids, 0, n, classes=np.arange(100000))
    q, _ = synth.make_device(Q, 512, seed_noise=synth.SEED_QUERY, n_classes=100000, noise=0.5, device=dev)
    for _ in range(2):
        clf.kneighbors_device(q)
    for _ in range(2):
        clf.kneighbors_device(q[:8].contiguous())
        clf.kneighbors_device(q[:1].contiguous())
    torch.cuda.synchronize()
