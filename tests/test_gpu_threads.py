"""Concurrency of the boundary: in the reference the mining generator runs on Keras' Sequence enqueuer THREAD beside the
training thread (tools/train.py:172-177: fit_generator(..., workers, use_multiprocessing=False)).
The C ABI promises re-entrancy per (thread, stream): per-thread scratch, no global mutable state, thread-local error
text.  Here a mining thread and a training thread hammer the library at the same time, each on its own CUDA stream,
and every result must equal the one computed serially."""
import threading

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


def _batches(n, classes, per, d, seed0):
    out = []
    for i in range(n):
        x, lab = synth.make_numpy(classes * per, d, seed_noise=seed0 + i, n_classes=classes, rows_per_class=per,
                                  noise=0.5, relu=True)
        out.append((unit_rows(x), lab.astype(np.int64)))
    return out


def test_mining_thread_beside_training_thread():
    from embeddingnet_b200 import losses_and_accuracies as lac
    from embeddingnet_b200.datagenerators import mine_batch_triplets
    from embeddingnet_b200.models import BankKNNClassifier

    mine_sets = _batches(12, 32, 8, 128, 100)            # BASELINE config 1 shaped sampled sets
    train_sets = _batches(12, 64, 8, 256, 200)           # 512 x 256 training batches
    bank, bl = synth.make_numpy(20_000, 128, n_classes=500, noise=0.5)
    clf = BankKNNClassifier(n_neighbors=5).fit(bank, bl)
    queries = [s[0][:64] for s in mine_sets]
    fn = lac.batch_hard_triplet_loss(0.5)
    ba = lac.batch_all_triplet_loss(0.5, max_positives=7)

    def mine_all():
        np.random.seed(5)                                # only this thread touches the global NumPy RNG
        res = []
        for (x, lab), q in zip(mine_sets, queries):
            trip, _ = mine_batch_triplets(x, lab, margin=0.5, mode="semihard")
            _, idx = clf.kneighbors(q)
            res.append((trip.copy(), idx.copy()))
        return res

    def train_all():
        res = []
        for x, lab in train_sets:
            e = torch.tensor(x, device="cuda", requires_grad=True)
            loss = fn(lab, e) + ba(lab, e)
            loss.backward()
            res.append((float(loss.item()), e.grad.cpu().numpy()))
        return res

    want_mine, want_train = mine_all(), train_all()
    torch.cuda.synchronize()
    got = {}
    errors = []

    def worker(name, body):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                for rep in range(3):                      # several rounds: more chances to interleave
                    got[name] = body()
                torch.cuda.current_stream().synchronize()
        except Exception as exc:                          # surfaced in the main thread below
            errors.append((name, exc))

    t = threading.Thread(target=worker, args=("mine", mine_all))
    t.start()
    worker("train", train_all)
    t.join()
    assert not errors, errors
    for (trip, idx), (wt, wi) in zip(got["mine"], want_mine):
        np.testing.assert_array_equal(trip, wt)
        np.testing.assert_array_equal(idx, wi)
    for (loss, g), (wl, wg) in zip(got["train"], want_train):
        assert loss == wl                                # the loss reductions run in a fixed order
        # batch-hard scatters its gradient with float atomics (several anchors may pick the same row): the sum is
        # the same up to the order of a handful of additions
        np.testing.assert_allclose(g, wg, rtol=0, atol=1e-6 * float(np.abs(wg).max()))
    # and against the oracle, once
    x, lab = train_sets[0]
    ref = float(O.batch_hard(lab, x, 0.5, False, False)["loss"]) + float(O.batch_all(lab, x, 0.5, False)["loss"])
    assert abs(want_train[0][0] - ref) <= 1e-5 * ref


def test_error_text_is_thread_local():
    """en_last_error() is per thread: an argument error raised on one thread does not leak into another."""
    import ctypes

    from embeddingnet_b200 import _lib

    lib = _lib.load()
    seen = {}

    def bad(name, total_len):
        rc = lib.en_triplet_apn_fwd(ctypes.c_void_p(16), 4, total_len, ctypes.c_float(0.5), ctypes.c_void_p(16), None)
        seen[name] = (rc, lib.en_last_error().decode())

    t = threading.Thread(target=bad, args=("a", 10))
    t.start()
    t.join()
    bad("b", 11)
    assert seen["a"][0] == -1 and "(10)" in seen["a"][1]
    assert seen["b"][0] == -1 and "(11)" in seen["b"][1]
