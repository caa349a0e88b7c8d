"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of EmbeddingNet's distance / mining / loss / bank-kNN path.

Nothing in the product package imports this module.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may use it, and only as the checker / baseline.

Parity status: the reference ships NO tests, golden vectors or fixtures for this path, and
TensorFlow 2.2 cannot run in this image.  The oracle is therefore pinned against *outputs of the reference's own
source files executed here under an import shim* (``oracle/ref_shim.py``; fixtures in ``tests/golden/`` made by
``tests/golden/make_golden.py``) plus the real scikit-learn the reference calls.  Rows that do not exist in the
reference at all (batch-hard / batch-all triplet, all-pairs contrastive -- BASELINE.json asks for them, the
reference only cites the papers) are "parity unpinned": their oracle is the published formula (Hermans et al. 2017,
and the TensorFlow formulation by O. Moindrot that the reference README cites, README.md:112,116), restated here in
float64.

All citations are ``file:line`` into /root/reference.
"""
from __future__ import annotations

import itertools

import numpy as np

F32 = np.float32


# ------------------------------------------------------------------------------------------------ losses (lac)
def contrastive_loss(y_true, y_pred):
    """embedding_net/losses_and_accuracies.py:4-11.  margin is the literal 1; returns the scalar mean (f32)."""
    y_true = np.asarray(y_true, F32)
    y_pred = np.asarray(y_pred, F32)
    margin = F32(1)
    square_pred = y_pred * y_pred
    margin_square = np.square(np.maximum(margin - y_pred, F32(0)))
    return np.mean(y_true * square_pred + (F32(1) - y_true) * margin_square, dtype=np.float64).astype(F32)


def contrastive_loss_grad(y_true, y_pred, upstream=1.0):
    """d contrastive_loss / d y_pred (what TF autodiff of lac:4-11 yields; relu' taken as (x > 0) on margin - d,
    and 0 contribution at exactly margin == d because the factor (margin - d) is 0 there)."""
    y_true = np.asarray(y_true, np.float64)
    y_pred = np.asarray(y_pred, np.float64)
    n = y_pred.size
    g = y_true * 2.0 * y_pred - (1.0 - y_true) * 2.0 * np.maximum(1.0 - y_pred, 0.0)
    return (g * (upstream / n)).astype(F32)


def triplet_loss(margin=0.5):
    """embedding_net/losses_and_accuracies.py:14-44: per-sample hinge on squared L2 over [a|p|n] thirds."""

    def loss_function(y_true, y_pred):
        y_pred = np.asarray(y_pred, F32)
        total = y_pred.shape[-1]
        a = y_pred[:, 0:int(total * 1 / 3)]                       # lac:29
        p = y_pred[:, int(total * 1 / 3):int(total * 2 / 3)]      # lac:30
        n = y_pred[:, int(total * 2 / 3):int(total * 3 / 3)]      # lac:31
        pos = np.sum(np.square(a - p), axis=1, dtype=np.float64)  # lac:34 (f64 accumulate, f32 result)
        neg = np.sum(np.square(a - n), axis=1, dtype=np.float64)  # lac:37
        basic = pos.astype(F32) - neg.astype(F32) + F32(margin)   # lac:40
        return np.maximum(basic, F32(0))                          # lac:41

    return loss_function


def triplet_loss_grad(y_pred, margin, upstream):
    """Gradient of sum_i upstream[i] * loss_function(...)[i] w.r.t. y_pred.  TF routes maximum(x, 0)'s gradient to
    x when x >= 0; an exactly-zero hinge argument has measure zero on the test inputs."""
    y = np.asarray(y_pred, np.float64)
    total = y.shape[-1]
    c1, c2, c3 = int(total * 1 / 3), int(total * 2 / 3), int(total * 3 / 3)
    a, p, n = y[:, :c1], y[:, c1:c2], y[:, c2:c3]
    basic = np.sum((a - p) ** 2, 1) - np.sum((a - n) ** 2, 1) + margin
    act = (basic >= 0).astype(np.float64) * np.asarray(upstream, np.float64).reshape(-1)
    g = np.zeros_like(y)
    g[:, :c1] = act[:, None] * (2 * (a - p) - 2 * (a - n))
    g[:, c1:c2] = act[:, None] * (-2 * (a - p))
    g[:, c2:c3] = act[:, None] * (2 * (a - n))
    return g.astype(F32)


def accuracy(y_true, y_pred):
    """embedding_net/losses_and_accuracies.py:47-50."""
    y_true = np.asarray(y_true, F32)
    y_pred = np.asarray(y_pred, F32)
    return np.mean(y_true == (y_pred < F32(0.5)).astype(y_true.dtype), dtype=np.float64).astype(F32)


def siamese_l2(e1, e2):
    """embedding_net/models.py:225: sqrt(max(sum((e1-e2)^2, axis=1, keepdims=True), K.epsilon()=1e-7))."""
    e1 = np.asarray(e1, np.float64)
    e2 = np.asarray(e2, np.float64)
    s = np.sum((e1 - e2) ** 2, axis=1, keepdims=True)
    return np.sqrt(np.maximum(s, 1e-7)).astype(F32)


def siamese_l2_grad(e1, e2, upstream):
    e1 = np.asarray(e1, np.float64)
    e2 = np.asarray(e2, np.float64)
    s = np.sum((e1 - e2) ** 2, axis=1, keepdims=True)
    d = np.sqrt(np.maximum(s, 1e-7))
    act = (s >= 1e-7).astype(np.float64)
    g1 = act * np.asarray(upstream, np.float64).reshape(-1, 1) * (e1 - e2) / d
    return g1.astype(F32), (-g1).astype(F32)


def siamese_l1(e1, e2):
    """embedding_net/models.py:218: abs(e1 - e2)."""
    return np.abs(np.asarray(e1, F32) - np.asarray(e2, F32))


def l2_normalize(x):
    """embedding_net/backbones.py:38,77,118: K.l2_normalize(x, axis=1) == x * rsqrt(max(sum x^2, 1e-12))."""
    x64 = np.asarray(x, np.float64)
    ss = np.sum(x64 * x64, axis=1, keepdims=True)
    return (x64 / np.sqrt(np.maximum(ss, 1e-12))).astype(F32)


def dense_relu(x, kernel, bias=None, normalize=False):
    """embedding_net/backbones.py:114-119: Dense(units, activation="relu") (+ K.l2_normalize) in float64."""
    y = np.asarray(x, np.float64) @ np.asarray(kernel, np.float64)
    if bias is not None:
        y = y + np.asarray(bias, np.float64)[None, :]
    y = np.maximum(y, 0.0)
    if normalize:
        ss = np.sum(y * y, axis=1, keepdims=True)
        y = y / np.sqrt(np.maximum(ss, 1e-12))
    return y.astype(F32)


def dense_relu_grad(x, kernel, bias, normalize, upstream):
    """float64 autograd gradients (gx, gw, gb) of ``dense_relu`` for an upstream gradient on its output
    (tf.nn.relu passes the gradient where the input is > 0; K.l2_normalize = x * rsqrt(max(sum x^2, 1e-12)))."""
    import torch

    xt = torch.tensor(np.asarray(x, np.float64), requires_grad=True)
    wt = torch.tensor(np.asarray(kernel, np.float64), requires_grad=True)
    bt = torch.tensor(np.asarray(bias, np.float64), requires_grad=True) if bias is not None else None
    y = xt @ wt
    if bt is not None:
        y = y + bt[None, :]
    y = torch.relu(y)
    if normalize:
        ss = (y * y).sum(dim=1, keepdim=True)
        y = y * torch.rsqrt(torch.clamp(ss, min=1e-12))
    y.backward(torch.tensor(np.asarray(upstream, np.float64)))
    return (xt.grad.numpy().astype(F32), wt.grad.numpy().astype(F32),
            bt.grad.numpy().astype(F32) if bt is not None else None)


def l2_normalize_grad(x, upstream):
    x = np.asarray(x, np.float64)
    g = np.asarray(upstream, np.float64)
    ss = np.sum(x * x, axis=1, keepdims=True)
    inv = 1.0 / np.sqrt(np.maximum(ss, 1e-12))
    dot = np.sum(g * x, axis=1, keepdims=True)
    gx = g * inv - np.where(ss >= 1e-12, x * dot * inv ** 3, 0.0)
    return gx.astype(F32)


# ------------------------------------------------------------------------------------------------ distances
def pairwise_distances_sklearn(x):
    """The reference's literal call, embedding_net/datagenerators.py:219 (scikit-learn, unpinned)."""
    from sklearn.metrics import pairwise_distances

    return pairwise_distances(np.asarray(x))


def pairwise_distances(x, squared=False):
    """Restatement of what sklearn does for float32 input (sklearn/metrics/pairwise.py, euclidean_distances +
    _euclidean_distances_upcast): float64 -2.x.y + |x|^2 + |y|^2, cast to float32, clamp at 0, zero diagonal,
    then sqrt in float32."""
    x64 = np.asarray(x, np.float64)
    xx = np.sum(x64 * x64, axis=1)
    d = -2.0 * (x64 @ x64.T)
    d += xx[:, None]
    d += xx[None, :]
    d = d.astype(F32)
    np.maximum(d, 0, out=d)
    np.fill_diagonal(d, 0)
    return d if squared else np.sqrt(d, out=d)


def sqdist_exact(a, b):
    """float64 direct sum((a-b)^2): the 'what is right' truth used to set tolerances and adjudicate near-ties."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    out = np.empty((a.shape[0], b.shape[0]), np.float64)
    step = max(1, int(2e7 // max(1, b.shape[0] * a.shape[1])))
    for s in range(0, a.shape[0], step):
        diff = a[s:s + step, None, :] - b[None, :, :]
        out[s:s + step] = np.einsum("ijk,ijk->ij", diff, diff)
    return out


# ------------------------------------------------------------------------------------------------ mining (dg)
def hardest_negative(loss_values, margin=0.5):
    """embedding_net/datagenerators.py:188-190."""
    hard_negative = np.argmax(loss_values)
    return hard_negative if loss_values[hard_negative] > 0 else None


def random_hard_negative(loss_values, margin=0.5):
    """embedding_net/datagenerators.py:192-194 (consumes the global legacy NumPy RNG)."""
    hard_negatives = np.where(loss_values > 0)[0]
    return np.random.choice(hard_negatives) if len(hard_negatives) > 0 else None


def semihard_negative(loss_values, margin=0.5):
    """embedding_net/datagenerators.py:196-199."""
    semihard_negatives = np.where(np.logical_and(loss_values < margin, loss_values > 0))[0]
    return np.random.choice(semihard_negatives) if len(semihard_negatives) > 0 else None


SELECTORS = {"semihard": semihard_negative, "hardest": hardest_negative, "random_hard": random_hard_negative}


def mine_batch_triplets(all_embeddings, k_classes, k_samples, margin, mode, distance_matrix=None):
    """Numeric core of TripletsDataGenerator.get_batch_triplets_mining, embedding_net/datagenerators.py:217-250.

    ``all_embeddings`` is the class-major (k_classes*k_samples, d) stack of dg:217.  Returns (triplets, used_fallback)
    with triplets an (T, 3) int array of (anchor, positive, negative) row ids in the reference's emission order.
    RNG: draws from the global ``np.random`` exactly as dg:194,199 do."""
    select = SELECTORS[mode]
    if distance_matrix is None:
        distance_matrix = pairwise_distances_sklearn(all_embeddings)        # dg:219
    n = k_classes * k_samples
    out = []
    anchor_positive = None
    negative_indices = None
    for idx in range(k_classes):                                              # dg:225
        current = np.zeros(n, dtype=bool)
        current[idx * k_samples:(idx + 1) * k_samples] = True               # dg:226-227
        positive_indices = np.where(current)[0]
        negative_indices = np.where(~current)[0]
        anchor_positives = np.array(list(itertools.combinations(positive_indices, 2)))   # dg:231
        ap_distances = distance_matrix[anchor_positives[:, 0], anchor_positives[:, 1]]   # dg:233
        for anchor_positive, ap_distance in zip(anchor_positives, ap_distances):
            loss_values = ap_distance - distance_matrix[anchor_positive[0], negative_indices] + margin  # dg:235
            loss_values = np.array(loss_values)
            hard = select(loss_values, margin=margin)                         # dg:237
            if hard is not None:
                out.append((anchor_positive[0], anchor_positive[1], negative_indices[hard]))  # dg:239-243
    fallback = False
    if len(out) == 0:                                                         # dg:246-250
        out.append((anchor_positive[0], anchor_positive[1], negative_indices[0]))
        fallback = True
    return np.array(out, dtype=np.int64), fallback


# ------------------------------------------------------------------------------------------------ in-batch losses (new API, D1)
def _dist_matrix64(emb, squared):
    """float64 distance matrix with the build's stated semantics: exact sum((a-b)^2), diagonal exactly 0,
    sqrt(0) = 0 (Moindrot's mask trick gives the same value and a zero gradient there)."""
    d2 = sqdist_exact(emb, emb)
    np.fill_diagonal(d2, 0.0)
    return d2 if squared else np.sqrt(d2)


def batch_hard(labels, emb, margin=0.5, squared=False, soft=False):
    """Batch-hard triplet loss (Hermans et al. 2017, eq. 5; Moindrot's batch_hard_triplet_loss).  Per anchor:
    hardest positive = max over same-label j != i (0 when there is none), hardest negative = min over other
    labels (row max when there is none, as in Moindrot's ``dist + rowmax * (1 - mask)``).  Ties -> lowest index.
    Returns dict(loss, per_anchor, hp, hn, hp_idx, hn_idx)."""
    labels = np.asarray(labels).reshape(-1)
    D = _dist_matrix64(emb, squared)
    B = D.shape[0]
    same = labels[:, None] == labels[None, :]
    eye = np.eye(B, dtype=bool)
    pos = same & ~eye
    neg = ~same
    hp = np.where(pos, D, 0.0).max(axis=1)
    hp_idx = np.where(pos.any(1), np.where(pos, D, -np.inf).argmax(axis=1), -1)
    rowmax = D.max(axis=1, keepdims=True)
    dn = D + rowmax * (~neg)
    hn = dn.min(axis=1)
    hn_idx = np.where(neg.any(1), np.where(neg, D, np.inf).argmin(axis=1), -1)
    z = hp - hn
    per = np.logaddexp(0.0, z) if soft else np.maximum(z + margin, 0.0)
    return dict(loss=F32(per.mean()), per_anchor=per.astype(F32), hp=hp.astype(F32), hn=hn.astype(F32),
                hp_idx=hp_idx.astype(np.int32), hn_idx=hn_idx.astype(np.int32))


def batch_all(labels, emb, margin=0.5, squared=False):
    """Batch-all triplet loss (Moindrot's batch_all_triplet_loss): sum of positive hinge terms over valid
    (i, j, k) divided by the number of terms > 1e-16.  Returns dict(loss, fraction, num_positive, num_valid).
    Computed anchor by anchor so B = 4096 stays tractable."""
    labels = np.asarray(labels).reshape(-1)
    D = _dist_matrix64(emb, squared)
    B = D.shape[0]
    total = 0.0
    num_pos = 0
    num_valid = 0
    for i in range(B):
        same = labels == labels[i]
        p = same.copy()
        p[i] = False
        n = ~same
        dp = D[i, p]
        dn = D[i, n]
        if dp.size == 0 or dn.size == 0:
            continue
        t = dp[:, None] - dn[None, :] + margin
        num_valid += t.size
        t = np.maximum(t, 0.0)
        total += t.sum()
        num_pos += int((t > 1e-16).sum())
    loss = total / (num_pos + 1e-16)
    return dict(loss=F32(loss), loss64=float(loss), fraction=F32(num_pos / (num_valid + 1e-16)),
                num_positive=num_pos, num_valid=num_valid)


def contrastive_allpairs(labels, emb):
    """All-pairs contrastive loss: lac:4-11 applied to every ordered pair i != j with y_ij = [label_i == label_j]
    and d_ij from the Siamese L2 head, models:225 (sqrt(max(d^2, 1e-7))).  Mean over the B*(B-1) pairs."""
    labels = np.asarray(labels).reshape(-1)
    d2 = sqdist_exact(emb, emb)
    np.fill_diagonal(d2, 0.0)
    d = np.sqrt(np.maximum(d2, 1e-7))
    B = d.shape[0]
    y = (labels[:, None] == labels[None, :]).astype(np.float64)
    t = y * d * d + (1.0 - y) * np.maximum(1.0 - d, 0.0) ** 2
    np.fill_diagonal(t, 0.0)
    return F32(t.sum() / (B * (B - 1)))


def _torch_grad(fn, emb):
    import torch

    e = torch.tensor(np.asarray(emb, np.float64), requires_grad=True)
    loss = fn(e)
    (g,) = torch.autograd.grad(loss, e)
    return loss.item(), g.numpy().astype(F32)


def _torch_dist(e, squared, eps_mask=True):
    import torch

    diff = e[:, None, :] - e[None, :, :]
    d2 = (diff * diff).sum(-1)
    if squared:
        return d2
    zero = d2 <= 0
    return torch.where(zero, torch.zeros_like(d2), torch.sqrt(torch.where(zero, torch.ones_like(d2), d2)))


def batch_hard_grad(labels, emb, margin=0.5, squared=False, soft=False):
    """float64 autograd gradient of ``batch_hard`` w.r.t. the embeddings (oracle for the fused backward)."""
    import torch

    lab = torch.tensor(np.asarray(labels).reshape(-1))

    def fn(e):
        D = _torch_dist(e, squared)
        same = lab[:, None] == lab[None, :]
        eye = torch.eye(len(lab), dtype=torch.bool)
        pos = same & ~eye
        neg = ~same
        hp = torch.where(pos, D, torch.zeros_like(D)).max(dim=1).values
        rowmax = D.max(dim=1, keepdim=True).values
        hn = (D + rowmax * (~neg)).min(dim=1).values
        z = hp - hn
        per = torch.nn.functional.softplus(z) if soft else torch.clamp(z + margin, min=0)
        return per.mean()

    return _torch_grad(fn, emb)


def batch_hard_grad_analytic(labels, emb, margin=0.5, squared=False, soft=False):
    """Closed-form float64 gradient of ``batch_hard`` from its own arg-max / arg-min (no B x B x d temporaries, so it
    runs at B = 4096, d = 512 where the autograd form would need 68 GB): per anchor
    g_i/B * ( s_p (e_i - e_p) - s_n (e_i - e_n) ), mirrored onto rows p and n; s = 2 (squared) or 1 / D (0 at D = 0);
    g_i = [hinge argument >= 0] (``torch.clamp`` passes the gradient at 0) or sigmoid(z) for the soft margin.
    Pinned to ``batch_hard_grad`` (autograd) by tests/test_oracle_cpu.py.  Returns (loss, grad float32)."""
    labels = np.asarray(labels).reshape(-1)
    e = np.asarray(emb, np.float64)
    B = e.shape[0]
    r = batch_hard(labels, emb, margin, squared, soft)
    D = _dist_matrix64(emb, squared)
    rows = np.arange(B)
    hp_idx, hn_idx = r["hp_idx"].astype(np.int64), r["hn_idx"].astype(np.int64)
    has_p, has_n = hp_idx >= 0, hn_idx >= 0
    if not has_n.all():  # no other label at all: Moindrot's row maximum stands in for the hardest negative
        hn_idx = np.where(has_n, hn_idx, np.where(np.eye(B, dtype=bool), -np.inf, D).argmax(axis=1))
        has_n = np.ones(B, bool) if B > 1 else has_n
    pi, ni = np.where(has_p, hp_idx, 0), np.where(has_n, hn_idx, 0)
    hp = np.where(has_p, D[rows, pi], 0.0)
    hn = np.where(has_n, D[rows, ni], 0.0)
    z = hp - hn
    g = (1.0 / (1.0 + np.exp(-z)) if soft else (z + margin >= 0).astype(np.float64)) / B
    if squared:
        sp, sn = 2.0 * g * has_p, 2.0 * g * has_n
    else:
        sp = np.where(has_p & (hp > 0), g / np.where(hp > 0, hp, 1.0), 0.0)
        sn = np.where(has_n & (hn > 0), g / np.where(hn > 0, hn, 1.0), 0.0)
    vp = sp[:, None] * (e - e[pi])
    vn = sn[:, None] * (e - e[ni])
    grad = vp - vn
    np.add.at(grad, pi, -vp)
    np.add.at(grad, ni, vn)
    return float(r["loss"]), grad.astype(F32)


def batch_all_grad(labels, emb, margin=0.5, squared=False):
    import torch

    lab = torch.tensor(np.asarray(labels).reshape(-1))

    def fn(e):
        D = _torch_dist(e, squared)
        B = len(lab)
        same = lab[:, None] == lab[None, :]
        eye = torch.eye(B, dtype=torch.bool)
        ap = (same & ~eye)[:, :, None]
        an = (~same)[:, None, :]
        t = D[:, :, None] - D[:, None, :] + margin
        t = torch.where(ap & an, t, torch.zeros_like(t)).clamp(min=0)
        npos = (t > 1e-16).sum()
        return t.sum() / (npos + 1e-16)

    return _torch_grad(fn, emb)


def contrastive_allpairs_grad(labels, emb):
    import torch

    lab = torch.tensor(np.asarray(labels).reshape(-1))

    def fn(e):
        B = len(lab)
        diff = e[:, None, :] - e[None, :, :]
        d2 = (diff * diff).sum(-1)
        d = torch.sqrt(torch.clamp(d2, min=1e-7))
        y = (lab[:, None] == lab[None, :]).to(e.dtype)
        t = y * d * d + (1 - y) * torch.clamp(1 - d, min=0) ** 2
        t = t * (1 - torch.eye(B, dtype=e.dtype))
        return t.sum() / (B * (B - 1))

    return _torch_grad(fn, emb)


def batch_all_grad_analytic(labels, emb, margin=0.5, squared=False):
    """Closed-form float64 gradient of ``batch_all`` (the count of positive triplets is piecewise constant, as in
    TF/torch autodiff): G_ap = +#{n active}/np, G_an = -#{p active}/np, grad = rowsum(W) E - W E with
    W = (G + G^T) * dD/d(d^2-ish factor).  Scales to B = 4096 (validated against autograd on small batches)."""
    labels = np.asarray(labels).reshape(-1)
    e = np.asarray(emb, np.float64)
    D = _dist_matrix64(e, squared)
    B = D.shape[0]
    G = np.zeros((B, B))
    npos = 0
    for i in range(B):
        same = labels == labels[i]
        p = np.where(same & (np.arange(B) != i))[0]
        n = np.where(~same)[0]
        if p.size == 0 or n.size == 0:
            continue
        act = (D[i, p][:, None] - D[i, n][None, :] + margin) > 1e-16
        npos += int(act.sum())
        G[i, p] += act.sum(axis=1)
        G[i, n] -= act.sum(axis=0)
    G /= (npos + 1e-16)
    with np.errstate(divide="ignore", invalid="ignore"):
        S = np.full((B, B), 2.0) if squared else np.where(D > 0, 1.0 / D, 0.0)
    W = (G + G.T) * S
    return (W.sum(axis=1, keepdims=True) * e - W @ e).astype(F32)


def contrastive_allpairs_grad_analytic(labels, emb):
    labels = np.asarray(labels).reshape(-1)
    e = np.asarray(emb, np.float64)
    d2 = sqdist_exact(e, e)
    np.fill_diagonal(d2, 0.0)
    B = d2.shape[0]
    d = np.sqrt(np.maximum(d2, 1e-7))
    y = labels[:, None] == labels[None, :]
    tp = np.where(y, 1.0, -np.maximum(1.0 - d, 0.0) / d) * (d2 >= 1e-7)
    np.fill_diagonal(tp, 0.0)
    W = 4.0 * tp / (B * (B - 1))
    return (W.sum(axis=1, keepdims=True) * e - W @ e).astype(F32)


# ------------------------------------------------------------------------------------------------ bank kNN (models)
def knn_exact(bank, queries, k, id_offset=0):
    """float64 brute-force k nearest neighbours ordered by (distance, id): the build's lowest-index tie rule
    (np.argmin at models:124 picks the first minimum).  Returns (dist f32 (Q,k), ids int64 (Q,k))."""
    bank = np.asarray(bank)
    queries = np.asarray(queries)
    Q = queries.shape[0]
    k = min(k, bank.shape[0])
    ids = np.empty((Q, k), np.int64)
    d2o = np.empty((Q, k), np.float64)
    step = max(1, int(4e7 // max(1, bank.shape[0])))
    b64 = bank.astype(np.float64)
    bb = np.sum(b64 * b64, axis=1)
    for s in range(0, Q, step):
        q64 = queries[s:s + step].astype(np.float64)
        # candidate generation with the GEMM form, exact re-evaluation of a generous candidate set
        approx = bb[None, :] - 2.0 * (q64 @ b64.T)
        kk = min(bank.shape[0], k + 16)
        cand = np.argpartition(approx, kk - 1, axis=1)[:, :kk] if kk < bank.shape[0] else np.tile(
            np.arange(bank.shape[0]), (q64.shape[0], 1))
        for r in range(q64.shape[0]):
            c = np.sort(cand[r])
            diff = b64[c] - q64[r]
            d2 = np.einsum("ij,ij->i", diff, diff)
            order = np.lexsort((c, d2))[:k]
            ids[s + r] = c[order] + id_offset
            d2o[s + r] = d2[order]
    return np.sqrt(d2o).astype(F32), ids


def knn_sklearn(bank, labels, queries, k):
    """What the reference's duck-typed classifier does (models:58,136,138): KNeighborsClassifier(brute)."""
    from sklearn.neighbors import KNeighborsClassifier

    clf = KNeighborsClassifier(n_neighbors=k, algorithm="brute")
    clf.fit(bank, labels)
    dist, ids = clf.kneighbors(queries, n_neighbors=k)
    return dist, ids, clf.predict(queries)


def knn_vote(neighbor_labels):
    """Uniform-weight majority vote of sklearn's KNeighborsClassifier.predict: the most frequent label among the
    k neighbours; a tie resolves to the smallest label in sorted order (``classes_`` is np.unique-sorted)."""
    neighbor_labels = np.asarray(neighbor_labels)
    out = []
    for row in neighbor_labels:
        vals, counts = np.unique(row, return_counts=True)
        out.append(vals[np.argmax(counts)])
    return np.array(out)


def knn_sharded(bank, queries, k, n_shards):
    """CPU emulation of the multi-GPU path: contiguous row shards, per-shard top-k, merge by (dist, id)."""
    N = bank.shape[0]
    per = (N + n_shards - 1) // n_shards
    parts_d, parts_i = [], []
    for r in range(n_shards):
        lo, hi = r * per, min(N, (r + 1) * per)
        if lo >= hi:
            continue
        d, i = knn_exact(bank[lo:hi], queries, k, id_offset=lo)
        parts_d.append(d.astype(np.float64) ** 2)
        parts_i.append(i)
    # merge on exact float64 distances recomputed from ids to avoid the f32 rounding of the per-shard lists
    ids = np.concatenate(parts_i, axis=1)
    out_i = np.empty((queries.shape[0], min(k, N)), np.int64)
    out_d = np.empty((queries.shape[0], min(k, N)), np.float64)
    b64 = bank.astype(np.float64)
    for q in range(queries.shape[0]):
        c = ids[q]
        diff = b64[c] - queries[q].astype(np.float64)
        d2 = np.einsum("ij,ij->i", diff, diff)
        order = np.lexsort((c, d2))[:out_i.shape[1]]
        out_i[q] = c[order]
        out_d[q] = d2[order]
    return np.sqrt(out_d).astype(F32), out_i


def predict_1nn(bank, bank_labels, encoding):
    """EmbeddingNet.predict, embedding_net/models.py:122-126: distances to every bank row, np.argmin -> label."""
    d2 = sqdist_exact(np.asarray(encoding).reshape(1, -1), bank)[0]
    return bank_labels[int(np.argmin(d2))]


def prediction_accuracy(bank, bank_labels, queries, query_labels, k):
    """EmbeddingNet.calculate_prediction_accuracy, embedding_net/models.py:144-161, with predict_knn
    (models:128-142): top1 = classifier.predict()[0] == label, top5 = label among the labels of the 5 nearest."""
    _, ids = knn_exact(bank, queries, max(k, 5))
    lab = np.asarray(bank_labels)
    pred = knn_vote(lab[ids[:, :k]])
    top5 = lab[ids[:, :5]]
    ql = np.asarray(query_labels)
    top1 = float(np.mean(pred == ql))
    t5 = float(np.mean([ql[i] in top5[i] for i in range(len(ql))]))
    return {"top1": top1, "top5": t5}


# ------------------------------------------------------------------------------------------------ bank mining (C4)
def mine_bank_hardest(bank, labels, anchors_idx, k=1):
    """Nearest *negatives* of each anchor row over the whole bank (generalisation of dg:188-190 to bank scale:
    the hardest negative of every (a, p) pair is the nearest other-class row of a).  (dist, id) ordering."""
    bank64 = np.asarray(bank, np.float64)
    labels = np.asarray(labels)
    ids = np.empty((len(anchors_idx), k), np.int64)
    dist = np.empty((len(anchors_idx), k), np.float64)
    for r, a in enumerate(anchors_idx):
        diff = bank64 - bank64[a]
        d2 = np.einsum("ij,ij->i", diff, diff)
        d2[labels == labels[a]] = np.inf
        order = np.lexsort((np.arange(len(d2)), d2))[:k]
        ids[r] = order
        dist[r] = d2[order]
    return np.sqrt(dist).astype(F32), ids


def mine_bank_modes(bank, labels, anchors, anchor_labels, pos_d, margin, mode):
    """datagenerators.py:188-199 over a whole bank: for each (anchor, positive slot) the candidates among bank rows of
    another class, loss = (d_ap - d_an) + margin in float32 (dg:235), d = sqrtf(float32(float64 sum (a-n)^2));
    np.random.choice(candidates) == candidates[randint(len)] in (anchor, slot) order; -1 where there is none.
    Returns (ids (A, S) int64, counts (A, S, 2) [random_hard, semihard])."""
    bank64 = np.asarray(bank, np.float64)
    labels = np.asarray(labels)
    anchors = np.asarray(anchors, F32)
    pos_d = np.asarray(pos_d, F32)
    A, S = pos_d.shape
    ids = np.full((A, S), -1, np.int64)
    counts = np.zeros((A, S, 2), np.int64)
    m = F32(margin)
    for i in range(A):
        diff = bank64 - anchors[i].astype(np.float64)
        dn = np.sqrt(np.einsum("ij,ij->i", diff, diff).astype(F32))            # float32 sqrt of the float32 value
        neg = np.where(labels != anchor_labels[i])[0]
        for s in range(S):
            if pos_d[i, s] < 0:
                continue
            loss = (pos_d[i, s] - dn[neg]) + m                                    # float32, left to right
            hard = neg[loss > 0]
            semi = neg[(loss > 0) & (loss < m)]
            counts[i, s] = (len(hard), len(semi))
            if mode == "hardest":
                if len(neg):
                    j = int(np.argmax(loss))
                    if loss[j] > 0:
                        ids[i, s] = neg[j]
                continue
            cand = hard if mode == "random_hard" else semi
            if len(cand):
                ids[i, s] = cand[np.random.randint(0, len(cand))]
    return ids, counts


# ------------------------------------------------------------------------------------------------ CPU baselines (bench.py)
def batch_hard_loss_grad_cpu(labels, emb, margin=0.5):
    """The headline step on the host, built from the reference's own distance call
    (sklearn.metrics.pairwise_distances, embedding_net/datagenerators.py:219; multi-threaded BLAS) + NumPy masks and
    the closed-form batch-hard gradient.  Used only as bench.py's reported CPU baseline."""
    labels = np.asarray(labels).reshape(-1)
    emb = np.asarray(emb, np.float32)
    D = pairwise_distances_sklearn(emb)
    B = D.shape[0]
    same = labels[:, None] == labels[None, :]
    pos = same.copy()
    np.fill_diagonal(pos, False)
    hp_idx = np.where(pos, D, -1.0).argmax(axis=1)
    hn_idx = np.where(same, np.inf, D).argmin(axis=1)
    r = np.arange(B)
    hp, hn = D[r, hp_idx], D[r, hn_idx]
    z = hp - hn + margin
    loss = float(np.maximum(z, 0).mean())
    act = (z >= 0).astype(np.float32) / B
    sp = np.where(hp > 0, act / np.maximum(hp, 1e-30), 0).astype(np.float32)[:, None]
    sn = np.where(hn > 0, act / np.maximum(hn, 1e-30), 0).astype(np.float32)[:, None]
    vp = sp * (emb - emb[hp_idx])
    vn = sn * (emb - emb[hn_idx])
    g = vp - vn
    np.add.at(g, hp_idx, -vp)
    np.add.at(g, hn_idx, vn)
    return loss, g
