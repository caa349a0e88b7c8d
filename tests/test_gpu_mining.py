"""GPU parity: pairwise distances and in-batch hard-negative mining vs scikit-learn, the oracle and the golden
outputs of the reference's own TripletsDataGenerator (bit-exact indices)."""
import numpy as np
import pytest
import torch

from conftest import MINING_CASES, class_tables, unit_rows
from embeddingnet_b200 import synth
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built(lib_built):
    return lib_built


def test_pairwise_exact_matches_sklearn_golden(golden):
    from embeddingnet_b200.datagenerators import pairwise_distances

    x, _ = synth.make_numpy(96, 64, n_classes=12, rows_per_class=8, noise=0.5)
    D = pairwise_distances(x)
    ref = golden["pairwise_96x64"]
    assert D.dtype == np.float32 and np.all(np.diag(D) == 0)
    assert np.max(np.abs(D - ref)) <= 2e-6 and np.mean(D == ref) > 0.99
    np.testing.assert_array_equal(D, D.T)


@pytest.mark.parametrize("n,d", [(1, 4), (5, 3), (60, 256), (256, 128), (300, 100), (1000, 77)])
@pytest.mark.parametrize("squared", [False, True])
def test_pairwise_exact_vs_oracle(n, d, squared):
    from embeddingnet_b200.datagenerators import pairwise_distances

    x, _ = synth.make_numpy(n, d, n_classes=max(1, n // 8), rows_per_class=8, noise=0.5)
    if n > 4:
        x[3] = x[1]  # duplicated sample (np.random.choice(replace=True), dg:205): off-diagonal ~0
    D = pairwise_distances(x, squared=squared)
    ref = O.pairwise_distances(x, squared=squared)
    scale = float(ref.max()) + 1e-6
    assert np.max(np.abs(D - ref)) <= 2e-6 * scale
    assert np.mean(D == ref) > 0.98


@pytest.mark.parametrize("n,d", [(256, 128), (300, 100), (4096, 512)])
def test_pairwise_tensor_core_path(n, d):
    from embeddingnet_b200.datagenerators import pairwise_distances

    x, _ = synth.make_numpy(n, d, n_classes=max(1, n // 8), rows_per_class=8, noise=0.5, relu=True)
    x = unit_rows(x)
    D2 = pairwise_distances(x, squared=True, exact=False)
    ref = O.pairwise_distances(x, squared=True)
    # 3xTF32 + norm expansion on unit vectors: absolute error ~1e-6 on d^2 in [0, 4]
    assert np.max(np.abs(D2 - ref)) < 2e-5
    assert np.all(np.diag(D2) == 0)


@pytest.mark.parametrize("case", list(MINING_CASES))
@pytest.mark.parametrize("mode", ["hardest", "semihard", "random_hard"])
@pytest.mark.parametrize("seed", [7, 8])
def test_generator_matches_reference_generator(golden, case, mode, seed):
    """Drop-in TripletsDataGenerator vs the reference's, same RNG seed: identical (a, p, n) image ids, and the
    legacy NumPy RNG left in the identical state."""
    from embeddingnet_b200.datagenerators import TripletsDataGenerator

    ncls, per, d, kc, ks, margin, norm = MINING_CASES[case]
    tables = class_tables(ncls, per, d, norm)
    table = np.vstack(tables)
    names = ["c%04d" % i for i in range(ncls)]
    files = {n: ["%d" % (i * per + j) for j in range(per)] for i, n in enumerate(names)}

    class Model:
        def predict(self, images):
            return table[np.asarray(images).reshape(-1).astype(np.int64)]

    class Gen(TripletsDataGenerator):
        def _get_images_set(self, clss, idxs, with_aug=True):
            return np.array([[int(self.class_files_paths[clss][i])] for i in idxs], dtype=np.int64)

    g = Gen(Model(), files, names, n_batches=1, k_classes=kc, k_samples=ks, margin=margin,
            negatives_selection_mode=mode)
    np.random.seed(seed)
    (A, P, N), targets = g[0]
    got = np.stack([A.ravel(), P.ravel(), N.ravel()], axis=1)
    want = golden["mine_%s_%s_%d" % (case, mode, seed)]
    np.testing.assert_array_equal(got, want)
    assert targets.shape == (want.shape[0],) and np.all(targets == 1)
    assert np.random.random_sample() == golden["mine_%s_%s_%d_rng_after" % (case, mode, seed)]


@pytest.mark.parametrize("mode", ["hardest", "semihard"])
def test_generator_device_resident_path(golden, mode):
    """image hook and embedding model return CUDA tensors -> embeddings, mining and the A/P/N gather
    (en_gather_triplet_rows) stay on the device; same triplets and RNG state as the reference generator."""
    import torch
    from embeddingnet_b200.datagenerators import TripletsDataGenerator, gather_triplets

    case, seed = list(MINING_CASES)[0], 7
    ncls, per, d, kc, ks, margin, norm = MINING_CASES[case]
    table = torch.tensor(np.vstack(class_tables(ncls, per, d, norm)), device="cuda")
    names = ["c%04d" % i for i in range(ncls)]
    files = {n: ["%d" % (i * per + j) for j in range(per)] for i, n in enumerate(names)}

    class Model:
        def predict(self, images):  # "images" carry their row id in element 0 and a payload behind it
            return table[images[:, 0, 0].to(torch.int64)]

    class Gen(TripletsDataGenerator):
        def _get_images_set(self, clss, idxs, with_aug=True):
            ids = torch.tensor([float(self.class_files_paths[clss][i]) for i in idxs], device="cuda")
            return torch.stack([ids, ids * 0.5, ids + 1.0, -ids, ids * 2.0], dim=1).reshape(-1, 5, 1)

    g = Gen(Model(), files, names, n_batches=1, k_classes=kc, k_samples=ks, margin=margin,
            negatives_selection_mode=mode)
    np.random.seed(seed)
    (A, P, N), targets = g[0]
    assert all(isinstance(t, torch.Tensor) and t.is_cuda and t.shape[1:] == (5, 1) for t in (A, P, N))
    got = np.stack([A[:, 0, 0].cpu().numpy(), P[:, 0, 0].cpu().numpy(), N[:, 0, 0].cpu().numpy()], axis=1)
    want = golden["mine_%s_%s_%d" % (case, mode, seed)]
    np.testing.assert_array_equal(got.astype(np.int64), want)
    np.testing.assert_array_equal(A[:, 4, 0].cpu().numpy(), want[:, 0] * 2.0)  # whole rows travel, not just the id
    assert np.random.random_sample() == golden["mine_%s_%s_%d_rng_after" % (case, mode, seed)]
    # gather kernel against torch indexing, odd row length (scalar copy path) and an empty triplet list
    rows = torch.arange(7 * 13, dtype=torch.float32, device="cuda").reshape(7, 13)
    trip = np.array([[0, 6, 3], [5, 5, 1], [2, 0, 6]], dtype=np.int64)
    a, p_, n_ = gather_triplets(rows, trip)
    for k, out in enumerate((a, p_, n_)):
        assert torch.equal(out, rows[torch.tensor(trip[:, k], device="cuda")])
    assert gather_triplets(rows, np.zeros((0, 3), np.int64))[0].shape == (0, 13)


@pytest.mark.parametrize("mode", ["hardest", "semihard", "random_hard"])
def test_mine_batch_vs_oracle_general_labels(mode):
    from embeddingnet_b200.datagenerators import mine_batch_triplets

    kc, ks, d = 13, 5, 40
    x, _ = synth.make_numpy(kc * ks, d, n_classes=kc, rows_per_class=ks, noise=0.6, relu=True)
    x = unit_rows(x)
    labels = np.repeat(np.arange(kc), ks)
    for margin in (0.05, 0.5, 2.5):
        np.random.seed(3)
        want, fb_w = O.mine_batch_triplets(x, kc, ks, margin, mode)
        s_w = np.random.random_sample()
        np.random.seed(3)
        got, fb_g = mine_batch_triplets(x, labels, margin=margin, mode=mode)
        s_g = np.random.random_sample()
        np.testing.assert_array_equal(got, want)
        assert fb_w == fb_g and s_w == s_g


def test_mining_fallback_when_no_triplet():
    """dg:246-250: nothing selected -> one fallback triplet (last pair, first negative)."""
    from embeddingnet_b200.datagenerators import mine_batch_triplets

    # two far-apart tight classes: every loss value is negative
    x = np.zeros((6, 4), np.float32)
    x[3:, 0] = 100.0
    x += (np.arange(24).reshape(6, 4) * 1e-3).astype(np.float32)
    labels = np.repeat(np.arange(2), 3)
    want, fb = O.mine_batch_triplets(x, 2, 3, 0.5, "hardest")
    got, fb2 = mine_batch_triplets(x, labels, margin=0.5, mode="hardest")
    assert fb and fb2
    np.testing.assert_array_equal(got, want)


def test_selection_callables_on_vectors(golden):
    from embeddingnet_b200.datagenerators import TripletsDataGenerator

    g = TripletsDataGenerator(None, {"a": ["0"]}, ["a"])
    lv, _ = synth.make_numpy(40, 33, seed_noise=31337)
    np.random.seed(123)
    res = []
    for r in range(lv.shape[0]):
        row = lv[r] * (0.2 if r % 4 == 0 else 1.0) - (0.9 if r % 5 == 0 else 0.0)
        out = [g.hardest_negative(row, margin=0.5), g.random_hard_negative(row, margin=0.5),
               g.semihard_negative(row, margin=0.5)]
        res.append([-1 if v is None else v for v in out])
    np.testing.assert_array_equal(np.asarray(res), golden["select_vectors"])
    with pytest.raises(KeyError):
        TripletsDataGenerator(None, {"a": ["0"]}, ["a"], negatives_selection_mode="nope")
