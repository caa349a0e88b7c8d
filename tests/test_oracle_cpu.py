"""CPU tests: the oracle (oracle/np_oracle.py) against the committed golden outputs of the reference's own source,
and -- where /root/reference is present -- against the reference executed live under the import shim."""
import numpy as np
import pytest

from conftest import MINING_CASES, class_tables
from embeddingnet_b200 import synth
from oracle import np_oracle as O
from oracle import ref_shim


@pytest.mark.parametrize("name,B,d", [("lac_small", 16, 32), ("lac_refcfg", 60, 256)])
def test_lac_oracle_matches_reference_outputs(golden, name, B, d):
    y, _ = synth.make_numpy(B, 3 * d, seed_noise=4242 + B)
    loss = O.triplet_loss(0.5)(None, y)
    # pos - neg cancels: the reference's own float32 result carries ~1 ulp of |a-p|^2 (up to ~170 here), so the
    # absolute tolerance is 1e-5 relative to that magnitude, not to the (much smaller) hinge value
    scale = float(np.max(np.sum((y[:, :d] - y[:, d:2 * d]) ** 2, axis=1)))
    np.testing.assert_allclose(loss, golden[name + "_triplet_loss"], rtol=1e-5, atol=1e-6 * scale)
    g = O.triplet_loss_grad(y, 0.5, golden[name + "_upstream"])
    np.testing.assert_allclose(g, golden[name + "_triplet_grad"], rtol=1e-4, atol=1e-6)
    dcol, _ = synth.make_numpy(B, 1, seed_noise=555 + B)
    dcol = np.abs(dcol) * 1.6
    yt = (np.arange(B) % 2).astype(np.float32).reshape(B, 1)
    np.testing.assert_allclose(O.contrastive_loss(yt, dcol), golden[name + "_contrastive_loss"], rtol=1e-5)
    np.testing.assert_allclose(O.contrastive_loss_grad(yt, dcol), golden[name + "_contrastive_grad"], rtol=1e-4,
                               atol=1e-7)
    assert O.accuracy(yt, dcol) == golden[name + "_accuracy"]


def test_pairwise_restatement_is_bit_exact_with_sklearn(golden):
    x, _ = synth.make_numpy(96, 64, n_classes=12, rows_per_class=8, noise=0.5)
    ours = O.pairwise_distances(x)
    ref = golden["pairwise_96x64"]
    assert ours.dtype == ref.dtype == np.float32
    # float64 GEMM summation order may differ from BLAS by one float32 ulp on a few entries
    assert np.max(np.abs(ours - ref)) <= 2e-6
    assert np.mean(ours == ref) > 0.99
    assert np.all(np.diag(ours) == 0)
    truth = np.sqrt(O.sqdist_exact(x, x))
    np.testing.assert_allclose(ours, truth, rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("case", list(MINING_CASES))
@pytest.mark.parametrize("mode", ["hardest", "semihard", "random_hard"])
@pytest.mark.parametrize("seed", [7, 8])
def test_mining_restatement_matches_reference_generator(golden, case, mode, seed):
    ncls, per, d, kc, ks, margin, norm = MINING_CASES[case]
    tables = class_tables(ncls, per, d, norm)
    # replay the reference's RNG draws (dg:202,205) to rebuild the batch, then run the restated core
    np.random.seed(seed)
    cls = np.random.choice(ncls, size=kc, replace=False)
    picks = [np.random.choice(per, size=ks, replace=True) for _ in cls]
    rows = np.concatenate([c * per + p for c, p in zip(cls, picks)])
    table = np.vstack(tables)
    trip, _ = O.mine_batch_triplets(table[rows], kc, ks, margin, mode)
    got = rows[trip]
    np.testing.assert_array_equal(got, golden["mine_%s_%s_%d" % (case, mode, seed)])
    assert np.random.random_sample() == golden["mine_%s_%s_%d_rng_after" % (case, mode, seed)]


def test_choice_equals_randint_stream():
    """np.random.choice(c) consumes the legacy stream exactly like c[np.random.randint(0, len(c))]."""
    for seed in range(50):
        c = np.arange(3, 3 + (seed % 17) + 1)
        np.random.seed(seed)
        a = [np.random.choice(c) for _ in range(5)]
        s1 = np.random.random_sample()
        np.random.seed(seed)
        b = [c[np.random.randint(0, len(c))] for _ in range(5)]
        s2 = np.random.random_sample()
        assert a == b and s1 == s2


def test_select_vectors(golden):
    lv, _ = synth.make_numpy(40, 33, seed_noise=31337)
    np.random.seed(123)
    res = []
    for r in range(lv.shape[0]):
        row = lv[r] * (0.2 if r % 4 == 0 else 1.0) - (0.9 if r % 5 == 0 else 0.0)
        out = [O.hardest_negative(row, 0.5), O.random_hard_negative(row, 0.5), O.semihard_negative(row, 0.5)]
        res.append([-1 if v is None else v for v in out])
    np.testing.assert_array_equal(np.asarray(res), golden["select_vectors"])


def test_knn_oracle_matches_sklearn(golden):
    bank, labels = synth.make_numpy(600, 48, n_classes=30, rows_per_class=0, noise=0.5)
    q, _ = synth.make_numpy(50, 48, seed_noise=synth.SEED_QUERY, n_classes=30, noise=0.6)
    dist, ids = O.knn_exact(bank, q, 5)
    np.testing.assert_array_equal(ids, golden["knn_idx"])          # tie-free inputs
    np.testing.assert_allclose(dist, golden["knn_dist"], rtol=1e-5)
    pred = O.knn_vote(labels[ids])
    np.testing.assert_array_equal(pred, golden["knn_pred"])


def test_knn_sharded_emulation_is_shard_invariant():
    bank, _ = synth.make_numpy(500, 24, n_classes=10, noise=0.5)
    bank[17] = bank[3]   # exact duplicates: ties must go to the lowest id for every shard count
    bank[250] = bank[3]
    q = bank[[3, 40, 499]] + np.float32(0.0)
    d1, i1 = O.knn_exact(bank, q, 5)
    assert i1[0, 0] == 3 and i1[0, 1] == 17 and i1[0, 2] == 250
    for shards in (2, 3, 4, 8):
        d, i = O.knn_sharded(bank, q, 5, shards)
        np.testing.assert_array_equal(i, i1)
        np.testing.assert_allclose(d, d1, rtol=1e-6)


def test_batch_hard_and_all_small_known_answer():
    # 4 points on a line, labels [0,0,1,1]: distances are exact small integers
    emb = np.array([[0.0], [1.0], [3.0], [7.0]], np.float32)
    lab = np.array([0, 0, 1, 1])
    bh = O.batch_hard(lab, emb, margin=0.5, squared=False)
    # anchors: 0: hp=1, hn=3 -> 0 ; 1: hp=1, hn=2 -> 0 ; 2: hp=4, hn=2 -> 2.5 ; 3: hp=4, hn=6 -> 0
    np.testing.assert_allclose(bh["per_anchor"], [0, 0, 2.5, 0])
    assert bh["hp_idx"].tolist() == [1, 0, 3, 2] and bh["hn_idx"].tolist() == [2, 2, 1, 1]
    ba = O.batch_all(lab, emb, margin=0.5, squared=False)
    # valid triplets (a,p,n): (0,1,2):-1.5 (0,1,3):-5.5 (1,0,2):-0.5 (1,0,3):-4.5 (2,3,0):1.5 (2,3,1):2.5 (3,2,0):-2.5 (3,2,1):-1.5
    assert ba["num_valid"] == 8 and ba["num_positive"] == 2
    np.testing.assert_allclose(ba["loss"], 2.0)
    l, g = O.batch_hard_grad(lab, emb, 0.5, False)
    np.testing.assert_allclose(l, 2.5 / 4)
    # only anchor 2 active: d(hp)/de2 = -1 (p=3 is to the right), d(-hn)/de2 = -1 (n=1 to the left) ...
    np.testing.assert_allclose(g.reshape(-1), np.array([0, 1, -2, 1]) / 4.0)


def test_analytic_gradients_match_autograd():
    from conftest import unit_rows

    x, lab = synth.make_numpy(60, 24, n_classes=10, rows_per_class=6, noise=0.5, relu=True)
    x = unit_rows(x)
    for squared in (False, True):
        _, g = O.batch_all_grad(lab, x, 0.5, squared)
        ga = O.batch_all_grad_analytic(lab, x, 0.5, squared)
        np.testing.assert_allclose(ga, g, rtol=1e-5, atol=1e-9)
    x7 = (x * 0.7).astype(np.float32)
    _, g = O.contrastive_allpairs_grad(lab, x7)
    np.testing.assert_allclose(O.contrastive_allpairs_grad_analytic(lab, x7), g, rtol=1e-5, atol=1e-10)


def test_batch_hard_analytic_gradient_matches_autograd():
    """The closed form used at the headline shape (B = 4096, d = 512) is pinned to the autograd oracle on small
    batches: duplicated rows (exact ties, zero distances), a single-class batch, squared / soft variants."""
    from conftest import unit_rows

    x, lab = synth.make_numpy(60, 24, n_classes=10, rows_per_class=6, noise=0.5, relu=True)
    x = unit_rows(x)
    x[7] = x[6]
    x[31] = x[2]
    for squared in (False, True):
        for soft in (False, True):
            l, g = O.batch_hard_grad(lab, x, 0.5, squared, soft)
            la, ga = O.batch_hard_grad_analytic(lab, x, 0.5, squared, soft)
            assert abs(l - la) <= 1e-6 * abs(l)
            np.testing.assert_allclose(ga, g, rtol=1e-5, atol=1e-9)
    x1, lab1 = synth.make_numpy(6, 5, n_classes=1, rows_per_class=6)
    l, g = O.batch_hard_grad(lab1, x1, 0.5)
    la, ga = O.batch_hard_grad_analytic(lab1, x1, 0.5)
    assert abs(l - la) <= 1e-6 * abs(l)
    np.testing.assert_allclose(ga, g, rtol=1e-5, atol=1e-9)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present (GPU box)")
def test_oracle_matches_reference_live():
    import torch

    lac, dg, K = ref_shim.load_reference()
    y, _ = synth.make_numpy(33, 3 * 20, seed_noise=77)
    ref = lac.triplet_loss(0.3)(None, K.wrap(torch.tensor(y))).numpy()
    np.testing.assert_allclose(O.triplet_loss(0.3)(None, y), ref, rtol=1e-5, atol=1e-6)
    # un-normalised mining batch through the reference generator vs the restatement, all modes
    tables = class_tables(10, 7, 16, True)
    for mode in ("hardest", "semihard", "random_hard"):
        g, table = ref_shim.make_reference_generator(tables, 6, 5, 0.4, mode)
        np.random.seed(5)
        (A, P, N), _ = g.get_batch_triplets_mining()
        np.random.seed(5)
        cls = np.random.choice(10, size=6, replace=False)
        picks = [np.random.choice(7, size=5, replace=True) for _ in cls]
        rows = np.concatenate([c * 7 + p for c, p in zip(cls, picks)])
        trip, _ = O.mine_batch_triplets(table[rows], 6, 5, 0.4, mode)
        np.testing.assert_array_equal(rows[trip], np.stack([A.ravel(), P.ravel(), N.ravel()], 1))


@pytest.mark.parametrize("mode", ["hardest", "random_hard", "semihard"])
def test_bank_mining_oracle_is_the_selection_callables_over_the_whole_bank(mode):
    """O.mine_bank_modes (the oracle of the bank-scale mining kernels) must be nothing but the reference's selection
    callables (dg:188-199, pinned above by the golden `select_vectors`) applied to loss_values = (d_ap - d_an) + m
    over all bank rows of another class, pairs in (anchor, slot) order, same RNG stream."""
    bank, labels = synth.make_numpy(400, 24, n_classes=11, noise=0.7, relu=True)
    a_idx = np.array([3, 57, 200, 399, 123])
    anchors, a_lab = bank[a_idx], labels[a_idx]
    rng = np.random.RandomState(4)
    pos_d = (1.2 + rng.rand(5, 3)).astype(np.float32)
    pos_d[2, 1] = -1.0  # unused slot
    margin = 0.6
    np.random.seed(9)
    got, counts = O.mine_bank_modes(bank, labels, anchors, a_lab, pos_d, margin, mode)
    after = np.random.random_sample()
    fn = {"hardest": O.hardest_negative, "random_hard": O.random_hard_negative, "semihard": O.semihard_negative}[mode]
    np.random.seed(9)
    want = np.full((5, 3), -1, np.int64)
    b64 = bank.astype(np.float64)
    for i in range(5):
        dn = np.sqrt(((b64 - anchors[i].astype(np.float64)) ** 2).sum(1).astype(np.float32))
        neg = np.flatnonzero(labels != a_lab[i])
        for s in range(3):
            if pos_d[i, s] < 0:
                continue
            loss_values = pos_d[i, s] - dn[neg] + margin                         # dg:235
            k = fn(loss_values, margin)
            if k is not None:
                want[i, s] = neg[k]                                              # dg:240
            assert counts[i, s, 0] == int((loss_values > 0).sum())
            assert counts[i, s, 1] == int(((loss_values > 0) & (loss_values < margin)).sum())
    np.testing.assert_array_equal(got, want)
    assert np.random.random_sample() == after
    assert (got >= 0).sum() >= 5


def test_dense_relu_oracle_known_answers():
    """backbones.py:114-119: Dense(relu) then K.l2_normalize; identity kernel, bias shift, all-negative row -> zeros."""
    x = np.array([[3.0, -4.0, 0.0], [-1.0, -2.0, -3.0], [0.0, 0.0, 2.0]], np.float32)
    eye = np.eye(3, dtype=np.float32)
    np.testing.assert_array_equal(O.dense_relu(x, eye), np.maximum(x, 0))
    y = O.dense_relu(x, eye, bias=np.array([0.0, 4.0, 0.0], np.float32), normalize=True)
    np.testing.assert_allclose(y[0], [1.0, 0.0, 0.0], atol=1e-7)                 # relu(3, 0, 0) normalised
    np.testing.assert_allclose(y[1], [0.0, 1.0, 0.0], atol=1e-7)                 # relu(-1, 2, -3) normalised
    assert np.all(O.dense_relu(x[1:2], eye, normalize=True) == 0)                # zero row stays zero (eps clamp)
    w, _ = synth.make_numpy(3, 5, seed_noise=3)
    np.testing.assert_allclose(O.dense_relu(x, w, normalize=True), O.l2_normalize(np.maximum(x @ w, 0)), rtol=1e-6,
                               atol=1e-7)
