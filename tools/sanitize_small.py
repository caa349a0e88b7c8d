"""Small-shape pass over the hot-path entry points for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python tools/sanitize_small.py

Shapes are tiny (the sanitizer slows kernels 10-100x) but cover every kernel family: row-wise, pairwise + in-batch
mining, the three fused tcgen05 losses (ragged sizes: partial tiles, out-of-bounds TMA rows), the bank scan in both
forms, the exact fallback, bank mining and the Dense head.  Results are checked against the oracle so that a clean
sanitizer log also means a correct run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from embeddingnet_b200 import losses_and_accuracies as lac, synth  # noqa: E402
from embeddingnet_b200.datagenerators import mine_batch_triplets  # noqa: E402
from embeddingnet_b200.models import BankKNNClassifier  # noqa: E402
from oracle import np_oracle as O  # noqa: E402


def unit(x):
    return (x / np.sqrt(np.maximum((x.astype(np.float64) ** 2).sum(1, keepdims=True), 1e-12))).astype(np.float32)


def main():
    dev = torch.device("cuda", 0)
    # row-wise
    x, lab = synth.make_numpy(37 * 5, 100, n_classes=37, rows_per_class=5, noise=0.5, relu=True)
    xn = lac.l2_normalize(torch.tensor(x, device=dev, requires_grad=True))
    xn.sum().backward()
    y = torch.tensor(np.concatenate([x, x[::-1], x * 0.5], axis=1), device=dev, requires_grad=True)
    lac.triplet_loss(0.5)(None, y).sum().backward()
    lac.siamese_l2_distance(torch.tensor(x, device=dev), torch.tensor(x[::-1].copy(), device=dev))
    # fused losses, ragged sizes (185 rows = 2 row tiles, d = 100)
    xu = unit(x)
    lab64 = lab.astype(np.int64)
    for fn, ref in ((lac.batch_hard_triplet_loss(0.5), float(O.batch_hard(lab64, xu, 0.5)["loss"])),
                    (lac.batch_all_triplet_loss(0.5, max_positives=4), float(O.batch_all(lab64, xu, 0.5)["loss"])),
                    (lac.contrastive_loss_all_pairs(), float(O.contrastive_allpairs(lab64, xu)))):
        e = torch.tensor(xu, device=dev, requires_grad=True)
        loss = fn(lab64, e)
        loss.backward()
        assert abs(loss.item() - ref) <= 1e-5 * abs(ref), (loss.item(), ref)
    # fast + slow batch-hard finalize (d = 128, duplicated rows saturate a slot)
    x2, lab2 = synth.make_numpy(64 * 8, 128, n_classes=64, rows_per_class=8, noise=0.5, relu=True)
    x2 = unit(x2)
    x2[9:12] = x2[300]
    e = torch.tensor(x2, device=dev, requires_grad=True)
    loss = lac.batch_hard_triplet_loss(0.5)(lab2.astype(np.int64), e)
    loss.backward()
    ref = float(O.batch_hard(lab2.astype(np.int64), x2, 0.5)["loss"])
    assert abs(loss.item() - ref) <= 1e-5 * ref
    # batch-all with large classes: lists of 24 slots walked eight at a time (pair_tc_kernel<..., kBig>), ragged tile
    x3, lab3 = synth.make_numpy(7 * 20, 72, n_classes=7, rows_per_class=20, noise=0.5, relu=True)
    x3 = unit(x3)
    e = torch.tensor(x3, device=dev, requires_grad=True)
    loss = lac.batch_all_triplet_loss(0.5, max_positives=19)(lab3.astype(np.int64), e)
    loss.backward()
    ref = float(O.batch_all(lab3.astype(np.int64), x3, 0.5)["loss"])
    assert abs(loss.item() - ref) <= 1e-5 * ref + 1e-7
    # host-buffer pipeline (en_bh_host_pipe_*): three steps through two slots
    from embeddingnet_b200.fused import BatchHardHostPipeline
    pipe = BatchHardHostPipeline(x2.shape[0], x2.shape[1], margin=0.5, depth=2)
    pin = BatchHardHostPipeline.pinned
    e_h, l_h = pin(x2.shape), pin((x2.shape[0],), torch.int32)
    e_h.copy_(torch.from_numpy(x2))
    l_h.copy_(torch.from_numpy(lab2.astype(np.int32)))
    outs = [(pin((1,)), pin(x2.shape)) for _ in range(3)]
    tickets = [pipe.submit(e_h, l_h, o[0], o[1]) for o in outs]
    for t in tickets:
        pipe.wait(t)
    ref = float(O.batch_hard(lab2.astype(np.int64), x2, 0.5)["loss"])
    assert all(abs(float(o[0][0]) - ref) <= 1e-5 * ref for o in outs)
    pipe.close()
    # in-batch mining (reference semantics)
    np.random.seed(0)
    want, _ = O.mine_batch_triplets(xu[:160], 32, 5, 0.5, "semihard")
    np.random.seed(0)
    trip, _ = mine_batch_triplets(xu[:160], lab[:160], margin=0.5, mode="semihard")
    assert np.array_equal(trip, want)
    # bank kNN: tensor scan, streaming scan, sharded merge on one device, mining strategies
    bank, bl = synth.make_numpy(3000, 100, n_classes=60, noise=0.5)
    q, _ = synth.make_numpy(70, 100, seed_noise=synth.SEED_QUERY, n_classes=60, noise=0.5)
    clf = BankKNNClassifier(n_neighbors=5).fit(bank, bl)
    _, ri = O.knn_exact(bank, q, 5)
    assert np.array_equal(clf.kneighbors(q)[1], ri)
    assert np.array_equal(clf.kneighbors(q[:2])[1], ri[:2])
    clf.predict(q)
    np.random.seed(1)
    clf.mine_negatives(bank[:64], bl[:64], positives=bank[60:124].reshape(64, 1, 100), margin=0.5, mode="semihard")
    clf.mine_negatives(bank[:64], bl[:64], positives=bank[60:124].reshape(64, 1, 100), margin=0.5, mode="hardest")
    torch.cuda.synchronize()
    print("sanitize_small: ok")


if __name__ == "__main__":
    main()
