"""The rule `decode_four_cells_first_max` (yolo_decode.cu) relies on, pinned on the CPU against the oracle
(reference arithmetic: scores = sigmoid(obj) * sigmoid(cls); max / first argmax over classes, yolo_head.py:282-285):

    the winning class can be found on the LOGITS -- maximum m, its first index, the largest different logit z below
    it -- whenever   (1 - sigmoid(m)) (m - z) >= 2^-19,  m has no NaN rival and the winning score is >= 1e-30;
    every other cell takes the reference loop.

The test restates the kernel's per-cell decision in numpy fp32 and checks, on random and adversarial logits, that the
cells it sends down the fast path have exactly the oracle's class index and score bits, and that the fast path is the
common case on realistic logits (so the kernel is what the benchmark measures, not its fallback)."""
import numpy as np
import torch

from oracle import detect_ref

F32 = np.float32


def fast_path_decision(obj_logit: np.ndarray, cls: np.ndarray):
    """cls [N, C] fp32 logits, obj_logit [N].  Returns (fast mask [N], index [N], score [N]) in the kernel's arithmetic."""
    n, c = cls.shape
    m = cls[:, 0].copy()
    m2 = np.full(n, -np.inf, F32)
    bi = np.zeros(n, np.int64)
    bad = np.isnan(m)
    with np.errstate(invalid="ignore", over="ignore"):
        for k in range(1, c):
            z = cls[:, k]
            bad |= np.isnan(z)
            ne = z != m
            m2 = np.where(ne, np.fmax(m2, np.fmin(m, z)), m2)
            bi = np.where(z > m, k, bi)
            m = np.fmax(m, z)
        sig = lambda v: (F32(1) / (F32(1) + np.exp(-v.astype(F32)).astype(F32))).astype(F32)
        obj = sig(obj_logit)
        best = (obj * sig(m)).astype(F32)
        margin = (F32(2.0 ** -19) * (F32(1) + np.exp(m).astype(F32))).astype(F32)
        fast = ~bad & (margin <= F32(3.0e38)) & ((m - m2) >= margin) & (best >= F32(1e-30))
    return fast, bi, best


def oracle_max(obj_logit: np.ndarray, cls: np.ndarray):
    pred = torch.zeros(1, 1, 1, cls.shape[0], 5 + cls.shape[1])
    pred[0, 0, 0, :, 4] = torch.from_numpy(obj_logit)
    pred[0, 0, 0, :, 5:] = torch.from_numpy(cls)
    out = detect_ref.yolo_decode(pred, torch.ones(1, 2))
    sc = out["scores"][0, 0, 0].numpy()
    # first argmax with NaN never beating a number except at class 0 (the `c == 0 || sc > best` loop of the kernels
    # = torch.max's CPU result on NaN-free rows)
    return sc


def first_argmax(sc: np.ndarray):
    idx = np.zeros(sc.shape[0], np.int64)
    best = sc[:, 0].copy()
    with np.errstate(invalid="ignore"):
        for k in range(1, sc.shape[1]):
            gt = sc[:, k] > best
            idx = np.where(gt, k, idx)
            best = np.where(gt, sc[:, k], best)
    return idx, best


def check(obj_logit, cls, min_fast_fraction=0.0):
    obj_logit, cls = obj_logit.astype(F32), cls.astype(F32)
    fast, bi, best = fast_path_decision(obj_logit, cls)
    sc = oracle_max(obj_logit, cls)
    want_i, want_s = first_argmax(sc)
    assert np.array_equal(bi[fast], want_i[fast])
    # numpy exp + division here, torch.sigmoid in the oracle: a few ulp apart (the GPU test holds the KERNEL to the all-scores kernel bit for bit)
    assert np.allclose(best[fast], want_s[fast], rtol=1e-6, atol=0)
    assert fast.mean() >= min_fast_fraction, fast.mean()
    return fast


def test_first_maximum_rule_on_random_logits():
    rng = np.random.default_rng(0)
    for scale in (0.5, 1.5, 4.0):
        cls = rng.standard_normal((20000, 80)) * scale
        obj = rng.standard_normal(20000) * 2
        check(obj, cls, min_fast_fraction=0.999 if scale < 4 else 0.9)
    # bf16-quantised logits (the model's head output): duplicates of the maximum are common and stay on the fast path
    cls = torch.from_numpy(rng.standard_normal((20000, 80)).astype(F32)).to(torch.bfloat16).float().numpy()
    fast = check(rng.standard_normal(20000), cls, min_fast_fraction=0.999)
    dup = (cls == cls.max(-1, keepdims=True)).sum(-1) > 1
    assert dup.sum() > 100 and fast[dup].all()


def test_first_maximum_rule_on_adversarial_logits():
    rng = np.random.default_rng(1)
    n, c = 4000, 80
    cases = [
        15.0 + 4.0 * rng.standard_normal((n, c)),                          # saturated sigmoids: different logits, equal scores
        1.0 + rng.integers(0, 4, (n, c)) * 2.0 ** -20,                     # a few ulp apart
        rng.integers(-2, 3, (n, c)).astype(np.float64),                    # many duplicates
        np.zeros((n, c)),
        -95.0 + rng.standard_normal((n, c)),                               # sigmoid in the denormals
        100.0 * rng.standard_normal((n, c)),
        np.where(rng.random((n, c)) < 0.02, np.inf, rng.standard_normal((n, c)) * 30),
        np.where(rng.random((n, c)) < 0.02, -np.inf, rng.standard_normal((n, c))),
        np.full((n, c), -np.inf),
    ]
    for cls in cases:
        for obj in (rng.standard_normal(n) * 2, np.full(n, -80.0), np.full(n, -110.0), np.full(n, 30.0)):
            check(obj, cls)
    # NaN logits / objectness never take the fast path
    cls = rng.standard_normal((n, c))
    cls[rng.random((n, c)) < 0.01] = np.nan
    fast, _, _ = fast_path_decision(rng.standard_normal(n).astype(F32), cls.astype(F32))
    assert not fast[np.isnan(cls).any(-1)].any()
    fast, _, _ = fast_path_decision(np.full(n, np.nan, F32), rng.standard_normal((n, c)).astype(F32))
    assert not fast.any()
