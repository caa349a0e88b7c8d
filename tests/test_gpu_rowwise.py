"""GPU parity: memory-bound row-wise kernels vs the oracle and the golden outputs of the reference's own code."""
import numpy as np
import pytest
import torch

from embeddingnet_b200 import synth
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built(lib_built):
    return lib_built


def cuda(x, grad=False):
    t = torch.tensor(x, device="cuda")
    t.requires_grad_(grad)
    return t


@pytest.mark.parametrize("name,B,d", [("lac_small", 16, 32), ("lac_refcfg", 60, 256)])
def test_lac_against_reference_outputs(golden, name, B, d):
    from embeddingnet_b200 import losses_and_accuracies as lac

    y, _ = synth.make_numpy(B, 3 * d, seed_noise=4242 + B)
    yp = cuda(y, True)
    loss = lac.triplet_loss(0.5)(None, yp)
    assert loss.shape == (B,)
    scale = float(np.max(np.sum((y[:, :d] - y[:, d:2 * d]) ** 2, axis=1)))
    np.testing.assert_allclose(loss.detach().cpu().numpy(), golden[name + "_triplet_loss"], rtol=1e-5,
                               atol=1e-6 * scale)
    (loss * cuda(golden[name + "_upstream"])).sum().backward()
    np.testing.assert_allclose(yp.grad.cpu().numpy(), golden[name + "_triplet_grad"], rtol=1e-4, atol=1e-5)
    dcol, _ = synth.make_numpy(B, 1, seed_noise=555 + B)
    dcol = np.abs(dcol) * 1.6
    yt = (np.arange(B) % 2).astype(np.float32).reshape(B, 1)
    dp = cuda(dcol, True)
    cl = lac.contrastive_loss(cuda(yt), dp)
    np.testing.assert_allclose(cl.item(), golden[name + "_contrastive_loss"], rtol=1e-5)
    cl.backward()
    np.testing.assert_allclose(dp.grad.cpu().numpy(), golden[name + "_contrastive_grad"], rtol=1e-4, atol=1e-7)
    assert lac.accuracy(cuda(yt), cuda(dcol)).item() == golden[name + "_accuracy"]


def test_triplet_loss_rejects_bad_width():
    from embeddingnet_b200 import losses_and_accuracies as lac

    with pytest.raises(ValueError):
        lac.triplet_loss(0.5)(None, torch.zeros(4, 10, device="cuda"))


@pytest.mark.parametrize("B,d", [(1, 1), (7, 5), (128, 256), (4096, 512), (33, 1000)])
def test_l2_normalize_fwd_bwd(B, d):
    from embeddingnet_b200 import losses_and_accuracies as lac

    x, _ = synth.make_numpy(B, d, seed_noise=11, relu=True)
    if B > 2:
        x[1] = 0  # post-ReLU rows can be all zero (bb:116-119) -> output 0
    up, _ = synth.make_numpy(B, d, seed_noise=12)
    xt = cuda(x, True)
    y = lac.l2_normalize(xt)
    np.testing.assert_allclose(y.detach().cpu().numpy(), O.l2_normalize(x), rtol=1e-5, atol=1e-7)
    (y * cuda(up)).sum().backward()
    ref = O.l2_normalize_grad(x, up)
    np.testing.assert_allclose(xt.grad.cpu().numpy(), ref, rtol=1e-4, atol=1e-4 * np.abs(ref).max() + 1e-7)


@pytest.mark.parametrize("B,d", [(1, 3), (60, 256), (513, 129)])
def test_siamese_heads(B, d):
    from embeddingnet_b200 import losses_and_accuracies as lac

    e1, _ = synth.make_numpy(B, d, seed_noise=21)
    e2, _ = synth.make_numpy(B, d, seed_noise=22)
    if B > 1:
        e2[0] = e1[0]  # clamp branch: sqrt(max(0, 1e-7))
    up, _ = synth.make_numpy(B, 1, seed_noise=23)
    a, b = cuda(e1, True), cuda(e2, True)
    dist = lac.siamese_l2_distance(a, b)
    assert dist.shape == (B, 1)
    np.testing.assert_allclose(dist.detach().cpu().numpy(), O.siamese_l2(e1, e2), rtol=1e-5)
    (dist * cuda(up)).sum().backward()
    g1, g2 = O.siamese_l2_grad(e1, e2, up)
    np.testing.assert_allclose(a.grad.cpu().numpy(), g1, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(b.grad.cpu().numpy(), g2, rtol=1e-4, atol=1e-6)
    l1 = lac.siamese_l1_distance(cuda(e1), cuda(e2))
    np.testing.assert_array_equal(l1.cpu().numpy(), O.siamese_l1(e1, e2))


def test_synth_device_matches_numpy():
    for kw in (dict(), dict(n_classes=10, rows_per_class=4, noise=0.5, relu=True), dict(n_classes=7, noise=0.25)):
        x, l = synth.make_numpy(100, 33, row_offset=5, **kw)
        xd, ld = synth.make_device(100, 33, row_offset=5, **kw)
        np.testing.assert_array_equal(xd.cpu().numpy(), x)
        if kw:
            np.testing.assert_array_equal(ld.cpu().numpy(), l)


@pytest.mark.parametrize("B,n_in,units", [(128, 512, 256), (300, 100, 130), (5, 33, 7), (1000, 256, 512)])
@pytest.mark.parametrize("normalize", [False, True])
def test_dense_relu_head(B, n_in, units, normalize):
    """Dense(relu) [+ l2_normalize] head (backbones.py:114-119) vs the float64 oracle; ragged sizes, zero rows."""
    from embeddingnet_b200.backbones import DenseReLU

    x, _ = synth.make_numpy(B, n_in, seed_noise=11)
    w, _ = synth.make_numpy(n_in, units, seed_noise=12)
    w = (w / np.sqrt(n_in)).astype(np.float32)
    b, _ = synth.make_numpy(1, units, seed_noise=13)
    b = (0.1 * b[0]).astype(np.float32)
    if B > 4:
        x[3] = -np.abs(x[3]) * 100.0  # drives the whole row through the ReLU to (almost surely) zero
    layer = DenseReLU(w, b, normalize=normalize)
    got = layer(x).cpu().numpy()
    want = O.dense_relu(x, w, b, normalize)
    scale = np.abs(want).max() + 1e-30
    assert np.abs(got - want).max() <= 1e-5 * scale
    if normalize:
        nz = np.linalg.norm(want, axis=1) > 0
        np.testing.assert_allclose(np.linalg.norm(got[nz], axis=1), 1.0, rtol=1e-5)
    assert not np.isnan(got).any()


@pytest.mark.parametrize("B,n_in,units", [(128, 512, 256), (300, 100, 130), (5, 33, 7), (1000, 256, 512)])
@pytest.mark.parametrize("normalize", [False, True])
def test_dense_relu_head_backward(B, n_in, units, normalize):
    """Training direction of the head (the reference fits through Dense(relu) -> Dense(relu) -> l2_normalize,
    backbones.py:114-119): gx, gw, gb of en_dense_relu_bwd vs float64 autograd, gradients 1e-4 (norm-wise)."""
    import torch
    from embeddingnet_b200.backbones import DenseReLU

    x, _ = synth.make_numpy(B, n_in, seed_noise=11)
    w, _ = synth.make_numpy(n_in, units, seed_noise=12)
    w = (w / np.sqrt(n_in)).astype(np.float32)
    b, _ = synth.make_numpy(1, units, seed_noise=13)
    b = (0.1 * b[0]).astype(np.float32)
    g, _ = synth.make_numpy(B, units, seed_noise=14)
    if B > 4:
        x[3] = -np.abs(x[3]) * 100.0           # an all-zero output row: the normalisation clamp branch
    layer = DenseReLU(w, b, normalize=normalize, trainable=True)
    xt = torch.tensor(x, device="cuda", requires_grad=True)
    y = layer(xt)
    y.backward(torch.tensor(g, device="cuda"))
    gx, gw, gb = O.dense_relu_grad(x, w, b, normalize, g)

    def rel(a, ref):
        return float(np.linalg.norm(a.astype(np.float64) - ref) / (np.linalg.norm(ref) + 1e-30))

    assert rel(xt.grad.cpu().numpy(), gx) < 1e-4
    assert rel(layer.kernel.grad.cpu().numpy(), gw) < 1e-4
    assert rel(layer.bias.grad.cpu().numpy(), gb) < 1e-4
    want = O.dense_relu(x, w, b, normalize)
    assert np.abs(y.detach().cpu().numpy() - want).max() <= 1e-5 * (np.abs(want).max() + 1e-30)


def test_training_step_through_head_and_batch_hard():
    """features -> EmbeddingHead (two Dense + l2_normalize) -> batch-hard loss -> backward: one SGD step lowers the
    loss, and the head's weight gradient matches float64 autograd through the same chain."""
    import torch
    from embeddingnet_b200 import losses_and_accuracies as lac
    from embeddingnet_b200.backbones import EmbeddingHead

    f, lab = synth.make_numpy(256, 96, n_classes=32, rows_per_class=8, noise=0.5, relu=True)
    k1, _ = synth.make_numpy(96, 64, seed_noise=22)
    k2, _ = synth.make_numpy(64, 128, seed_noise=23)
    k1, k2 = (k1 / 10).astype(np.float32), (k2 / 8).astype(np.float32)
    b1, b2 = np.full(64, 0.05, np.float32), np.full(128, 0.02, np.float32)
    head = EmbeddingHead(k1, b1, k2, b2, trainable=True)
    fn = lac.batch_hard_triplet_loss(0.5)
    ft = torch.tensor(f, device="cuda")
    loss0 = fn(lab, head(ft))
    loss0.backward()
    # float64 reference of d loss / d k2 through the selected pairs (the selection is piecewise constant)
    emb = O.dense_relu(O.dense_relu(f, k1, b1), k2, b2, normalize=True)
    _, gemb = O.batch_hard_grad(lab.astype(np.int64), emb, 0.5)
    _, gw2, _ = O.dense_relu_grad(O.dense_relu(f, k1, b1), k2, b2, True, gemb)
    got = head.fc2.kernel.grad.cpu().numpy()
    assert np.linalg.norm(got - gw2) <= 2e-4 * np.linalg.norm(gw2)
    with torch.no_grad():
        for p_ in head.parameters():
            p_ -= 0.5 * p_.grad
    loss1 = fn(lab, head(ft))
    assert loss1.item() < loss0.item()


def test_embedding_head_feeds_the_bank_path_on_device():
    """EmbeddingHead.predict keeps CUDA tensors on the device and matches the two-layer oracle."""
    import torch
    from embeddingnet_b200.backbones import EmbeddingHead

    f, _ = synth.make_numpy(257, 96, seed_noise=21, relu=True)
    k1, _ = synth.make_numpy(96, 64, seed_noise=22)
    k2, _ = synth.make_numpy(64, 128, seed_noise=23)
    k1, k2 = (k1 / 10).astype(np.float32), (k2 / 8).astype(np.float32)
    b1 = np.full(64, 0.05, np.float32)
    b2 = np.full(128, -0.02, np.float32)
    head = EmbeddingHead(k1, b1, k2, b2, embeddings_normalization=True)
    want = O.dense_relu(O.dense_relu(f, k1, b1), k2, b2, normalize=True)
    got = head.predict(f)
    assert isinstance(got, np.ndarray) and np.abs(got - want).max() <= 2e-5
    got_d = head.predict(torch.tensor(f, device="cuda"))
    assert isinstance(got_d, torch.Tensor) and got_d.is_cuda
    assert np.abs(got_d.cpu().numpy() - want).max() <= 2e-5
