"""GPU parity: encoding-bank nearest neighbours (tensor-core scan + exact re-rank, streaming scan, merge, vote,
accuracy) vs scikit-learn's golden outputs and the float64 oracle.  Neighbour ids and labels are bit-exact."""
import numpy as np
import pytest
import torch

from conftest import unit_rows
from embeddingnet_b200 import synth
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built(lib_built):
    return lib_built


def test_knn_matches_sklearn_golden(golden):
    from embeddingnet_b200.models import BankKNNClassifier

    bank, labels = synth.make_numpy(600, 48, n_classes=30, rows_per_class=0, noise=0.5)
    q, _ = synth.make_numpy(50, 48, seed_noise=synth.SEED_QUERY, n_classes=30, noise=0.6)
    clf = BankKNNClassifier(n_neighbors=5).fit(bank, labels)
    dist, idx = clf.kneighbors(q, n_neighbors=5)
    np.testing.assert_array_equal(idx, golden["knn_idx"])
    np.testing.assert_allclose(dist, golden["knn_dist"], rtol=1e-5)
    np.testing.assert_array_equal(clf.predict(q), golden["knn_pred"])
    # the streaming (<= 8 queries) path must agree with the tensor-core path
    d1, i1 = clf.kneighbors(q[:3], n_neighbors=5)
    np.testing.assert_array_equal(i1, golden["knn_idx"][:3])
    np.testing.assert_allclose(d1, golden["knn_dist"][:3], rtol=1e-5)


@pytest.mark.parametrize("precision", ["bf16x3", "tf32x3"])
@pytest.mark.parametrize("N,d,Q,k", [(1000, 64, 200, 5), (5000, 512, 300, 5), (777, 100, 129, 1), (3000, 256, 130, 10),
                                     (130, 32, 17, 29), (4, 8, 9, 5)])
def test_knn_vs_oracle(N, d, Q, k, precision):
    from embeddingnet_b200.models import BankKNNClassifier

    bank, labels = synth.make_numpy(N, d, n_classes=max(2, N // 20), noise=0.5, relu=True)
    bank = unit_rows(bank)
    q, _ = synth.make_numpy(Q, d, seed_noise=synth.SEED_QUERY, n_classes=max(2, N // 20), noise=0.6, relu=True)
    q = unit_rows(q)
    clf = BankKNNClassifier(n_neighbors=min(k, 5), precision=precision).fit(bank, labels)
    if k > N:
        # scikit-learn raises here; padding with id -1 would make models.py:139's labels[idx] return the LAST label
        with pytest.raises(ValueError, match="n_neighbors <= n_samples_fit"):
            clf.kneighbors(q, n_neighbors=k)
        k = N
    dist, idx = clf.kneighbors(q, n_neighbors=k)
    assert clf.last_uncertified <= max(1, Q // 20)  # the certificate covers (nearly) every query on tie-free data
    rd, ri = O.knn_exact(bank, q, k)
    kk = ri.shape[1]
    np.testing.assert_array_equal(idx[:, :kk], ri)
    np.testing.assert_allclose(dist[:, :kk], rd, rtol=1e-5, atol=1e-7)
    if kk < k:
        assert np.all(idx[:, kk:] == -1)


@pytest.mark.parametrize("N,d,Q,k", [(777, 100, 5, 5), (20000, 512, 8, 5), (20000, 512, 16, 5), (5000, 64, 17, 1),
                                     (20000, 256, 33, 5), (40000, 512, 64, 5), (300, 32, 64, 3), (129, 512, 40, 5)])
def test_knn_small_query_sets_take_the_bank_stationary_scan(N, d, Q, k):
    """5 <= Q <= 64 (between models.py:122's one image per call and the batched accuracy loop): the bank-stationary
    tcgen05 scan (en_knn_smallq_topk) -- ids bit-exact vs the float64 oracle, ties to the lowest id, ragged bank
    tiles, and the same answer as the engine path."""
    from embeddingnet_b200 import _lib
    from embeddingnet_b200.models import BankKNNClassifier

    bank, labels = synth.make_numpy(N, d, n_classes=max(2, N // 20), noise=0.5, relu=True)
    bank = unit_rows(bank)
    bank[N // 2] = bank[3]                      # exact duplicates: equal distances, the lower id must come first
    bank[N - 1] = bank[3]
    q, _ = synth.make_numpy(Q, d, seed_noise=synth.SEED_QUERY, n_classes=max(2, N // 20), noise=0.6, relu=True)
    q = unit_rows(q)
    q[0] = bank[3]
    clf = BankKNNClassifier(n_neighbors=k).fit(bank, labels)
    clf.smallq_max_q = 64                       # also exercise the 64-wide instantiation (default cut-over: 32)
    assert _lib.load().en_ws_bytes_knn_smallq(Q, N, d, k) > 0
    dist, idx = clf.kneighbors(q, n_neighbors=k)
    rd, ri = O.knn_exact(bank, q, k)
    np.testing.assert_array_equal(idx, ri)
    np.testing.assert_allclose(dist, rd, rtol=1e-5, atol=1e-7)
    assert idx[0, 0] == 3 and (k < 3 or list(idx[0, :3]) == [3, N // 2, N - 1])
    # the engine path gives the same ids
    big = BankKNNClassifier(n_neighbors=k, precision="tf32x3").fit(bank, labels)   # no BF16 planes: engine path
    np.testing.assert_array_equal(big.kneighbors(q, n_neighbors=k)[1], ri)


@pytest.mark.parametrize("precision", ["bf16x3", "tf32x3"])
def test_knn_ties_resolve_to_lowest_id(precision):
    from embeddingnet_b200.models import BankKNNClassifier

    bank, labels = synth.make_numpy(2000, 64, n_classes=50, noise=0.5)
    for dup in (10, 700, 1500, 1999):
        bank[dup] = bank[5]
    q = np.concatenate([bank[[5, 300]], bank[[5]] + np.float32(1e-3)]).astype(np.float32)
    q = np.tile(q, (40, 1))  # > 8 queries -> tensor-core path
    clf = BankKNNClassifier(n_neighbors=5, precision=precision).fit(bank, labels)
    _, idx = clf.kneighbors(q, n_neighbors=5)
    _, ri = O.knn_exact(bank, q, 5)
    np.testing.assert_array_equal(idx, ri)
    assert idx[0].tolist() == [5, 10, 700, 1500, 1999]
    _, idx_s = clf.kneighbors(q[:3], n_neighbors=5)  # streaming path
    np.testing.assert_array_equal(idx_s, ri[:3])


@pytest.mark.parametrize("precision", ["bf16x3", "tf32x3"])
def test_certificate_flags_near_ties_and_exact_path_resolves_them(precision):
    """Twenty bank rows that differ from one another by ~1e-7 relative: far inside the scan's rounding error, so the
    tensor-core pass cannot order them and its k + 3 candidates need not contain the true top-5.  The certificate
    must refuse those queries (uncertified > 0) and the float64 brute force must return the oracle's ids."""
    from embeddingnet_b200.models import BankKNNClassifier

    rng = np.random.RandomState(7)
    bank, labels = synth.make_numpy(3000, 128, n_classes=30, noise=0.5)
    base = bank[17].copy()
    rows = rng.choice(np.arange(100, 3000), size=20, replace=False)
    for r in rows:
        bank[r] = base + (rng.randn(128) * 2e-7 * np.abs(base)).astype(np.float32)
    q = np.tile(base[None, :], (24, 1)).astype(np.float32)
    q += (rng.randn(24, 128) * 1e-3).astype(np.float32)
    other, _ = synth.make_numpy(40, 128, seed_noise=synth.SEED_QUERY, n_classes=30, noise=0.5)
    q = np.concatenate([q, other]).astype(np.float32)
    _, ri = O.knn_exact(bank, q, 5)
    clf = BankKNNClassifier(n_neighbors=5, precision=precision).fit(bank, labels)
    _, idx = clf.kneighbors(q)
    assert clf.last_uncertified >= 20  # the near-tie queries; the unrelated ones are certified
    assert clf.last_uncertified <= 30
    np.testing.assert_array_equal(idx, ri)
    _, idx_s = clf.kneighbors(q[:8])  # streaming path: fp32 arithmetic, same certificate logic
    np.testing.assert_array_equal(idx_s, ri[:8])
    # with certification off the call still runs (no host read) and the far-from-tie queries are right
    raw = BankKNNClassifier(n_neighbors=5, precision=precision, certify=False).fit(bank, labels)
    _, idx_r = raw.kneighbors(q)
    far = 24 + np.flatnonzero(np.arange(40) % 30 != 17)  # class 17 holds the near-duplicates
    np.testing.assert_array_equal(idx_r[far], ri[far])


@pytest.mark.parametrize("N,d,Q,k,excl", [(2000, 64, 7, 5, False), (1500, 100, 64, 3, True), (300, 512, 33, 29, False),
                                          (3, 16, 5, 5, False)])
def test_exact_brute_force_kernel(N, d, Q, k, excl):
    """en_knn_exact_topk (the path behind the certificate) against the float64 oracle, incl. duplicated rows."""
    from embeddingnet_b200 import _lib
    from embeddingnet_b200._runtime import ptr, stream_ptr

    lib = _lib.load()
    bank, labels = synth.make_numpy(N, d, n_classes=max(2, N // 50), noise=0.5)
    if N > 100:
        bank[N - 1] = bank[3]
        bank[N // 2] = bank[3]
    q, _ = synth.make_numpy(Q, d, seed_noise=synth.SEED_QUERY, n_classes=max(2, N // 50), noise=0.5)
    q[0] = bank[min(3, N - 1)]
    anchors = np.arange(Q) % N
    if excl:
        q = bank[anchors].copy()
    dev = torch.device("cuda")
    tb, tq = torch.tensor(bank, device=dev), torch.tensor(q, device=dev)
    tl = torch.tensor(labels.astype(np.int32), device=dev)
    tql = torch.tensor(labels[anchors].astype(np.int32), device=dev) if excl else None
    d2 = torch.empty((Q, k), dtype=torch.float64, device=dev)
    ids = torch.empty((Q, k), dtype=torch.int64, device=dev)
    ws = torch.empty(lib.en_ws_bytes_knn_exact(Q, N, d, k), dtype=torch.uint8, device=dev)
    _lib.call("en_knn_exact_topk", ptr(tq), Q, d, ptr(tb), N, 1000, k, ptr(tql), ptr(tl) if excl else None,
              ptr(d2), ptr(ids), ptr(ws), ws.numel(), stream_ptr())
    got = ids.cpu().numpy()
    if excl:
        rd, ri = O.mine_bank_hardest(bank, labels, anchors, k=k)
    else:
        rd, ri = O.knn_exact(bank, q, k)
    kk = ri.shape[1]
    np.testing.assert_array_equal(got[:, :kk], ri + 1000)
    np.testing.assert_allclose(np.sqrt(d2.cpu().numpy()[:, :kk]), rd, rtol=1e-6, atol=1e-7)
    if kk < k:
        assert np.all(got[:, kk:] == -1)


def test_predict_vote_and_accuracy():
    from embeddingnet_b200.models import BankKNNClassifier

    bank, labels = synth.make_numpy(3000, 64, n_classes=40, noise=0.9)
    q, ql = synth.make_numpy(500, 64, seed_noise=synth.SEED_QUERY, n_classes=40, noise=0.9)
    names = np.array(["class_%02d" % l for l in labels])  # reference labels are class-name strings
    clf = BankKNNClassifier(n_neighbors=5).fit(bank, names)
    _, ri = O.knn_exact(bank, q, 5)
    want = O.knn_vote(names[ri])
    np.testing.assert_array_equal(clf.predict(q), want)
    acc = clf.score_topk(q, np.array(["class_%02d" % l for l in ql]))
    ref = O.prediction_accuracy(bank, names, q, np.array(["class_%02d" % l for l in ql]), 5)
    assert acc == pytest.approx(ref)


def test_merge_kernel_is_shard_invariant():
    """1 / 2 / 3 / 8 shards on one GPU through en_knn_shard_topk + en_knn_merge: ids identical to the unsharded
    result (the multi-GPU path minus the all-gather)."""
    import ctypes
    from embeddingnet_b200 import _lib
    from embeddingnet_b200._runtime import ptr, stream_ptr
    from embeddingnet_b200.models import BankKNNClassifier

    bank, labels = synth.make_numpy(4001, 96, n_classes=60, noise=0.5)
    bank[3000] = bank[12]
    q, _ = synth.make_numpy(150, 96, seed_noise=synth.SEED_QUERY, n_classes=60, noise=0.5)
    q[0] = bank[12]
    k = 5
    base = BankKNNClassifier(n_neighbors=k).fit(bank, labels)
    d0, i0 = base.kneighbors(q)
    _, ri = O.knn_exact(bank, q, k)
    np.testing.assert_array_equal(i0, ri)
    lab_ids = np.unique(labels, return_inverse=True)[1].astype(np.int32)
    for world in (2, 3, 8):
        parts_d, parts_i = [], []
        for r in range(world):
            lo, hi = BankKNNClassifier.shard_bounds(len(bank), world, r)
            c = BankKNNClassifier(n_neighbors=k).fit_shard(bank[lo:hi], lab_ids, lo, len(bank))
            d2, ids = c._search(torch.tensor(q, device="cuda"), k)
            parts_d.append(d2)
            parts_i.append(ids)
        D = torch.stack(parts_d).contiguous()
        I = torch.stack(parts_i).contiguous()
        d2m = torch.empty_like(parts_d[0])
        idm = torch.empty_like(parts_i[0])
        _lib.call("en_knn_merge", ptr(D), ptr(I), world, q.shape[0], k, ptr(d2m), ptr(idm), stream_ptr())
        np.testing.assert_array_equal(idm.cpu().numpy(), i0)
        np.testing.assert_allclose(np.sqrt(d2m.cpu().numpy()), d0, rtol=1e-6)


def test_bank_mining_excludes_same_label():
    """Offline hard-negative mining over a bank (BASELINE config 4): nearest rows of OTHER classes."""
    from embeddingnet_b200.models import BankKNNClassifier

    bank, labels = synth.make_numpy(2500, 64, n_classes=25, rows_per_class=100, noise=0.5, relu=True)
    bank = unit_rows(bank)
    clf = BankKNNClassifier(n_neighbors=3).fit(bank, labels)
    anchors = np.arange(0, 2500, 7)
    dist, ids = clf.kneighbors_device(bank[anchors], n_neighbors=3,
                                      exclude_labels=torch.tensor(labels[anchors], dtype=torch.int32))
    rd, ri = O.mine_bank_hardest(bank, labels, anchors, k=3)
    np.testing.assert_array_equal(ids.cpu().numpy(), ri)
    np.testing.assert_allclose(dist.cpu().numpy(), rd, rtol=1e-5)


def test_embeddingnet_predict_api(tmp_path):
    """EmbeddingNet.predict / predict_knn / calculate_prediction_accuracy on encodings (models.py:115-161)."""
    from embeddingnet_b200.models import EmbeddingNet, BankKNNClassifier

    bank, labels = synth.make_numpy(400, 32, n_classes=20, noise=0.4)
    names = ["sign_%02d" % l for l in labels]
    net = EmbeddingNet({"model": {"input_shape": [48, 48, 3]}, "encodings": {"knn_k": 5}})
    net.encoded_training_data = {"paths": ["%d.png" % i for i in range(400)], "labels": names, "encodings": bank}
    net.save_encodings(net.encoded_training_data, str(tmp_path), "enc.pkl")
    net2 = EmbeddingNet({"model": {"input_shape": [48, 48, 3]}, "encodings": {"knn_k": 5}})
    net2.load_encodings(str(tmp_path / "enc.pkl"))
    q, _ = synth.make_numpy(6, 32, seed_noise=synth.SEED_QUERY, n_classes=20, noise=0.4)
    for i in range(6):
        want = O.predict_1nn(bank, names, q[i])
        assert net2.predict_encoding(q[i:i + 1]) == want
        dists = net2.calculate_distances(q[i:i + 1])
        assert dists.shape == (400,) and names[int(np.argmin(dists))] == want
        pred, top5 = net2.predict_knn_encoding(q[i:i + 1], with_top5=True)
        _, ri = O.knn_exact(bank, q[i:i + 1], 5)
        assert top5 == [names[j] for j in ri[0]]
        assert pred.shape == (1,) and pred[0] == O.knn_vote(np.array(names)[ri])[0]


def test_saved_bank_keeps_the_reference_layout_after_predict(tmp_path):
    """generate_encodings -> predict -> save_encodings (the order tools/train.py and test.py use): the pickle holds
    exactly the reference's {paths, labels, encodings} (models.py:80-90), no cached GPU classifier, and loads
    without this package; the 1-NN cache follows a replaced bank."""
    import pickle
    from embeddingnet_b200.models import EmbeddingNet

    bank, labels = synth.make_numpy(300, 32, n_classes=15, noise=0.4)
    names = ["c%02d" % l for l in labels]
    net = EmbeddingNet({"model": {}, "encodings": {"knn_k": 5}})
    net.encoded_training_data = {"paths": ["%d.png" % i for i in range(300)], "labels": names, "encodings": bank}
    net.fit_knn()
    q, _ = synth.make_numpy(3, 32, seed_noise=synth.SEED_QUERY, n_classes=15, noise=0.4)
    assert net.predict_encoding(q[:1]) == O.predict_1nn(bank, names, q[0])
    net.save_encodings(net.encoded_training_data, str(tmp_path), "enc.pkl")
    with open(tmp_path / "enc.pkl", "rb") as f:
        raw = pickle.load(f)
    assert sorted(raw) == ["encodings", "labels", "paths"]
    assert isinstance(raw["encodings"], np.ndarray) and raw["encodings"].dtype == np.float32
    assert sorted(k for k in net.encoded_training_data if not k.startswith("knn")) == ["encodings", "labels", "paths"]
    # replace the bank in place: predictions must follow it (no stale cache)
    bank2 = bank[::-1].copy()
    net.encoded_training_data["encodings"] = bank2
    net.encoded_training_data["labels"] = names[::-1]
    for i in range(3):
        assert net.predict_encoding(q[i:i + 1]) == O.predict_1nn(bank2, names[::-1], q[i])
    # calculate_distances: plain Euclidean distances over the fitted device bank; an exact match reports 0
    dists = net.calculate_distances(bank2[17])
    want = np.sqrt(O.sqdist_exact(bank2[17:18], bank2)[0])
    assert dists[17] == 0.0 and int(np.argmin(dists)) == 17
    np.testing.assert_allclose(dists, want, rtol=1e-5, atol=1e-6)


def test_calculate_distances_at_bank_scale():
    """1M x 256 bank (BASELINE config 4's size): one pass over the resident bank, no per-call upload or broadcast."""
    from embeddingnet_b200.models import EmbeddingNet

    n, d = 1_000_000, 256
    bank_t, _ = synth.make_device(n, d, n_classes=10_000, noise=0.5)
    bank = bank_t.cpu().numpy()
    net = EmbeddingNet({"model": {}, "encodings": {}})
    net.encoded_training_data = {"paths": [], "labels": list(range(n)), "encodings": bank}
    q = bank[123_457] + 0.01
    dists = net.calculate_distances(q)
    torch.cuda.synchronize()
    before = torch.cuda.memory_allocated()
    dists = net.calculate_distances(q)
    assert torch.cuda.memory_allocated() - before < 64 * 1024 * 1024   # nothing bank-sized is allocated per call
    assert dists.shape == (n,) and int(np.argmin(dists)) == 123_457
    rows = np.arange(0, n, 9973)
    want = np.sqrt(O.sqdist_exact(q[None], bank[rows])[0])
    np.testing.assert_allclose(dists[rows], want, rtol=1e-5)
    assert net.predict_encoding(q) == 123_457


def test_sharded_knn_over_nccl():
    """The only partitioned path: bank rows sharded over >= 2 GPUs, NCCL all-gather + merge (skipped on 1 GPU)."""
    import os
    import subprocess
    import sys

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dist_knn_worker.py")
    port = 29600 + os.getpid() % 300
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(port), worker],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "dist knn ok" in r.stdout


def _mining_case(n=6000, d=64, n_classes=40, A=150, S=3, normalize=True):
    bank, labels = synth.make_numpy(n, d, n_classes=n_classes, noise=0.7, relu=True)
    if normalize:
        bank = unit_rows(bank)
    rng = np.random.RandomState(5)
    a_idx = rng.choice(n, size=A, replace=False)
    anchors = bank[a_idx].copy()
    a_lab = labels[a_idx].astype(np.int32)
    pos = np.zeros((A, S, d), np.float32)
    for i, (r, l) in enumerate(zip(a_idx, a_lab)):
        same = np.flatnonzero((labels == l) & (np.arange(n) != r))
        pos[i] = bank[rng.choice(same, size=S, replace=False)]
    pos_d = np.sqrt(((anchors[:, None, :].astype(np.float64) - pos.astype(np.float64)) ** 2).sum(-1).astype(np.float32))
    return bank, labels.astype(np.int32), anchors, a_lab, pos, pos_d


@pytest.mark.parametrize("precision", ["bf16x3", "tf32x3"])
@pytest.mark.parametrize("mode", ["semihard", "random_hard", "hardest"])
def test_bank_mining_strategies_match_the_oracle(mode, precision):
    """datagenerators.py:188-199 over a bank (BASELINE config 4): candidate counts, the drawn candidate and the RNG
    stream equal the float64 oracle's, for both operand formats."""
    from embeddingnet_b200.models import BankKNNClassifier

    bank, labels, anchors, a_lab, pos, pos_d = _mining_case()
    clf = BankKNNClassifier(n_neighbors=1, precision=precision).fit_shard(bank, labels, 0, len(bank))
    np.random.seed(11)
    want, _ = O.mine_bank_modes(bank, labels, anchors, a_lab, pos_d, 0.5, mode)
    rng_after = np.random.random_sample()
    np.random.seed(11)
    got = clf.mine_negatives(anchors, a_lab, positives=pos, margin=0.5, mode=mode)
    np.testing.assert_array_equal(got, want)
    assert np.random.random_sample() == rng_after
    assert (want >= 0).mean() > 0.3  # the case is not vacuous


def test_bank_mining_hardest_float32_ties_resolve_like_argmax():
    """dg:188-190: np.argmax over the float32 loss.  Two other-class rows whose distances to the anchor differ in
    float64 but round to the same float32 loss: the reference picks the LOWER index, not the float64-nearer row."""
    from embeddingnet_b200.models import BankKNNClassifier

    rs = np.random.RandomState(5)
    d = 64
    bank = rs.rand(400, d).astype(np.float32)
    labels = (np.arange(400) % 20).astype(np.int32)
    anchors = bank[[3, 50, 111]].copy()
    a_lab = labels[[3, 50, 111]].copy()
    for a_i, (near, lower) in enumerate(((300, 7), (390, 12), (201, 30))):
        # `near` (higher id) sits very close to the anchor; `lower` (lower id) a hair farther: equal in float32
        step = rs.randn(d).astype(np.float32)
        step *= 0.05 / np.linalg.norm(step)
        bank[near] = anchors[a_i] + step
        bank[lower] = anchors[a_i] + step * np.float32(1.0 + 2e-8)
        bank[lower, 0] = np.nextafter(bank[lower, 0], np.float32(10.0)) if bank[lower, 0] == bank[near, 0] else bank[lower, 0]
        for r in (near, lower):
            if labels[r] == a_lab[a_i]:
                labels[r] = (labels[r] + 1) % 20
    pos = (anchors + 0.3 * rs.randn(3, d).astype(np.float32))[:, None, :]
    pos_d = np.sqrt(((anchors[:, None, :].astype(np.float64) - pos.astype(np.float64)) ** 2).sum(-1).astype(np.float32))
    want, _ = O.mine_bank_modes(bank, labels, anchors, a_lab, pos_d, 0.5, "hardest")
    clf = BankKNNClassifier(n_neighbors=1).fit_shard(bank, labels, 0, len(bank))
    got = clf.mine_negatives(anchors, a_lab, positives=pos, margin=0.5, mode="hardest")
    np.testing.assert_array_equal(got, want)
    # the case really contains a float32 tie that the float64 order would resolve the other way
    d64 = ((bank[None, :, :].astype(np.float64) - anchors[:, None, :].astype(np.float64)) ** 2).sum(-1)
    flips = 0
    for i in range(3):
        neg = np.where(labels != a_lab[i])[0]
        flips += int(neg[np.argmin(d64[i, neg])] != want[i, 0])
    assert flips >= 1


def test_bank_mining_counts_and_sharding_invariance():
    """Counts through the C ABI vs the oracle (ragged sizes, unused slots, un-normalised rows), and the two-shard
    protocol of SURVEY 8(e) emulated on one GPU: per-shard counts -> owner shard resolves the rank."""
    from embeddingnet_b200 import _lib
    from embeddingnet_b200._runtime import ptr, stream_ptr
    from embeddingnet_b200.models import BankKNNClassifier
    import ctypes

    bank, labels, anchors, a_lab, pos, pos_d = _mining_case(n=3001, d=100, n_classes=23, A=77, S=2, normalize=False)
    pos_d[5, 1] = -1.0  # unused slot
    _, want_counts = O.mine_bank_modes(bank, labels, anchors, a_lab, pos_d, 0.8, "semihard")
    lib = _lib.load()
    MS = _lib.EN_MINE_MAX_SLOTS
    dev = torch.device("cuda")
    A, d = anchors.shape
    pd = torch.full((A, MS), -1.0, device=dev)
    pd[:, :2] = torch.tensor(pos_d, device=dev)
    ta, tl = torch.tensor(anchors, device=dev), torch.tensor(a_lab, device=dev)
    parts = []
    for lo, hi in ((0, 1500), (1500, 3001)):
        c = BankKNNClassifier(n_neighbors=1).fit_shard(bank[lo:hi], labels, lo, len(bank))
        counts = torch.zeros((A, MS, 2), dtype=torch.int32, device=dev)
        ws = torch.empty(lib.en_ws_bytes_mine_bank(A, d), dtype=torch.uint8, device=dev)
        bl = c._labels[lo:hi].contiguous()
        _lib.call("en_mine_bank_count", ptr(ta), ptr(tl), ptr(pd), A, d, 2, ctypes.c_float(0.8), ptr(c._bank), ptr(c._hi),
                  ptr(c._lo), ptr(c._norms), ptr(bl), hi - lo, c._prec, ptr(counts), ptr(ws), ws.numel(), stream_ptr())
        parts.append((c, counts, ws, bl, lo, hi))
    total = sum(p[1].cpu().numpy().astype(np.int64) for p in parts)
    np.testing.assert_array_equal(total[:, :2, :], want_counts)
    assert np.all(total[:, 2:, :] == 0)
    # resolve a fixed rank (the middle candidate) through the owner shard; compare with the oracle's candidate list
    for col, mode in ((0, "random_hard"), (1, "semihard")):
        r = total[:, :, col] // 2
        sel = np.full((A, MS), -1, np.int64)
        first = parts[0][1].cpu().numpy()[:, :, col].astype(np.int64)
        for k, (c, counts, ws, bl, lo, hi) in enumerate(parts):
            local = np.where(total[:, :, col] > 0, r - (first if k == 1 else 0), -1)
            mine = counts.cpu().numpy()[:, :, col]
            local = np.where((local >= 0) & (local < mine), local, -1).astype(np.int32)
            out = torch.full((A, MS), -1, dtype=torch.int64, device=dev)
            _lib.call("en_mine_bank_select", ptr(ta), ptr(tl), ptr(pd), A, d, 2, ctypes.c_float(0.8),
                      _lib.EN_MODE_RANDOM_HARD if mode == "random_hard" else _lib.EN_MODE_SEMIHARD,
                      ptr(torch.tensor(local, device=dev)), ptr(c._bank), ptr(c._hi), ptr(c._lo), ptr(c._norms), ptr(bl),
                      hi - lo, lo, c._prec, ptr(out), ptr(ws), ws.numel(), stream_ptr())
            sel = np.maximum(sel, out.cpu().numpy())
        # oracle: the r-th candidate in ascending id
        b64 = bank.astype(np.float64)
        for i in range(A):
            dn = np.sqrt((((b64 - anchors[i].astype(np.float64)) ** 2).sum(1)).astype(np.float32))
            neg = np.flatnonzero(labels != a_lab[i])
            for s in range(2):
                if pos_d[i, s] < 0:
                    assert sel[i, s] == -1
                    continue
                loss = (np.float32(pos_d[i, s]) - dn[neg]) + np.float32(0.8)
                cand = neg[loss > 0] if mode == "random_hard" else neg[(loss > 0) & (loss < np.float32(0.8))]
                assert sel[i, s] == (cand[len(cand) // 2] if len(cand) else -1), (i, s, mode)


def test_knn_large_bank_planted_neighbours_and_empty_inputs():
    """Size-independent properties at a bank too large for the float64 oracle (2M x 128): every query is a bank row
    plus a small perturbation, so its nearest neighbour is known by construction; the result must not depend on the
    operand format or on how the bank is cut into shards, and (nearly) every query must carry a certificate."""
    import ctypes
    from embeddingnet_b200 import _lib
    from embeddingnet_b200._runtime import ptr, stream_ptr
    from embeddingnet_b200.models import BankKNNClassifier

    dev = torch.device("cuda")
    n, d, Q, k = 2_000_000, 128, 3000, 5
    bank, _ = synth.make_device(n, d, n_classes=20_000, noise=0.5, device=dev)
    lab = (torch.arange(n, device=dev) % 20_000).to(torch.int32)
    src = torch.arange(0, n, n // Q, device=dev)[:Q]
    noise, _ = synth.make_device(Q, d, seed_noise=99, device=dev)
    q = (bank[src] + 0.02 * noise).contiguous()
    results = {}
    for prec in ("bf16x3", "tf32x3"):
        clf = BankKNNClassifier(n_neighbors=k, precision=prec, device=dev).fit_shard(bank, lab, 0, n)
        dist, ids = clf.kneighbors_device(q)
        assert clf.last_uncertified <= Q // 100
        assert torch.equal(ids[:, 0], src)                                   # the planted row
        assert bool((dist[:, 1:] >= dist[:, :-1]).all())                     # ascending
        results[prec] = ids.cpu().numpy()
        if prec == "bf16x3":
            # three unequal shards through the C-ABI merge
            parts_d, parts_i = [], []
            for lo, hi in ((0, 700_001), (700_001, 1_500_000), (1_500_000, n)):
                c = BankKNNClassifier(n_neighbors=k, device=dev).fit_shard(bank[lo:hi], lab, lo, n)
                d2, ii = c._search(q, k)
                parts_d.append(d2)
                parts_i.append(ii)
                del c
            D, I = torch.stack(parts_d).contiguous(), torch.stack(parts_i).contiguous()
            d2m, idm = torch.empty_like(parts_d[0]), torch.empty_like(parts_i[0])
            _lib.call("en_knn_merge", ptr(D), ptr(I), 3, Q, k, ptr(d2m), ptr(idm), stream_ptr())
            np.testing.assert_array_equal(idm.cpu().numpy(), results[prec])
        # empty inputs
        dd, ii = clf.kneighbors(np.zeros((0, d), np.float32))
        assert dd.shape == (0, k) and ii.shape == (0, k)
        assert clf.predict(np.zeros((0, d), np.float32)).shape == (0,)
        assert clf.mine_negatives(np.zeros((0, d), np.float32), np.zeros(0, np.int32),
                                  pos_dist=np.zeros((0, 2), np.float32)).shape == (0, 2)
        del clf
    np.testing.assert_array_equal(results["bf16x3"], results["tf32x3"])


@pytest.mark.parametrize("nranks", [1, 2])
def test_sharded_knn_through_the_c_abi_alone(nranks, tmp_path):
    """The exchange step of the sharded path through en_comm_* (NCCL bound by the library, no torch.distributed):
    one process per GPU, unique id handed over through a file, packed all-gather + en_knn_merge_packed, and the
    all-reduce(max) of the mining protocol.  nranks = 1 runs on the single-GPU box too."""
    import os
    import subprocess
    import sys

    if torch.cuda.device_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "comm_abi_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), str(nranks), str(tmp_path / "id.bin"),
                               str(tmp_path / ("out%d.npz" % r))], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(nranks)]
    outs = [p.communicate(timeout=300)[0].decode(errors="replace") for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)[-3000:]
    N, d, Q, k = 6000, 96, 150, 5
    bank, _ = synth.make_numpy(N, d, n_classes=60, noise=0.5)
    q, _ = synth.make_numpy(Q, d, seed_noise=synth.SEED_QUERY, n_classes=60, noise=0.5)
    rd, ri = O.knn_exact(bank, q, k)
    want_red = np.arange(Q) + 1000 * (np.arange(Q) % nranks)
    for r in range(nranks):
        z = np.load(tmp_path / ("out%d.npz" % r))
        np.testing.assert_array_equal(z["ids"], ri)
        np.testing.assert_allclose(np.sqrt(z["d2"]), rd, rtol=1e-5)
        np.testing.assert_array_equal(z["reduced"], want_red)
