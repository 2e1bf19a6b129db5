"""The walk `nms_sorted_kernel` (nms.cu) runs over the sorted candidates, restated in numpy and pinned on the CPU against
the oracle's greedy loops (reference: YOLODetectionHead.non_max_suppression yolo_head.py:678-731 and NMSFilter._standard_nms
postprocessing.py:505-607):

    candidates in descending score (ties -> lower index), CHUNK at a time: every candidate is first tested against the boxes
    kept in EARLIER chunks; then the chunk's groups of GROUP candidates are resolved in order -- (A) the group's candidates
    against the boxes kept earlier in this chunk and against each other (the GROUP x GROUP pair matrix as bit masks),
    (B) greedy over the group's alive word with those masks; the walk stops at max_det.

Small CHUNK / GROUP values push a few hundred candidates through many chunks and groups, which the kernel's 1024 / 32 only
does for thousands.  A candidate's fate depends only on the kept boxes before it, so the restatement must give exactly the
oracle's keep list -- in both suppression rules (agnostic: keep iff iou < thr; class-aware: suppress iff same class and
iou > thr), with ties, NaN boxes and caps."""
import numpy as np
import pytest

from oracle import detect_ref

F32 = np.float32


def suppresses(kb, cb, kcls, ccls, class_aware, thr):
    iou = detect_ref.iou_one_to_many(kb, cb[None])[0]
    if class_aware:
        return bool(kcls == ccls and iou > thr)
    return not bool(iou < thr)


def walk(boxes, scores, classes, thr, max_det, class_aware, chunk, group, score_thr=None):
    thr = F32(thr)
    idx = np.arange(len(scores))
    if score_thr is not None:
        idx = idx[scores > F32(score_thr)]                          # order-preserving compaction (NaN never passes)
    order = idx[np.lexsort((idx, -scores[idx].astype(np.float64)))]   # the stable sort: ties -> lower index
    kept = []                                                       # indices into the input
    for base in range(0, len(order), chunk):
        if len(kept) >= max_det:
            break
        kept0 = len(kept)
        cand = order[base:base + chunk]
        alive = np.array([not any(suppresses(boxes[k], boxes[c], classes[k], classes[c], class_aware, thr) for k in kept[:kept0])
                          for c in cand])
        for g0 in range(0, len(cand), group):
            if len(kept) >= max_det:
                break
            grp = cand[g0:g0 + group]
            aw = alive[g0:g0 + group].copy()
            if not aw.any():
                continue
            # step A: against the boxes kept earlier in this chunk, and the pair matrix of the group
            for j, c in enumerate(grp):
                if aw[j] and any(suppresses(boxes[k], boxes[c], classes[k], classes[c], class_aware, thr) for k in kept[kept0:]):
                    aw[j] = False
            pair = np.zeros((len(grp), len(grp)), bool)
            for i in range(len(grp)):
                for j in range(i + 1, len(grp)):
                    pair[i, j] = suppresses(boxes[grp[i]], boxes[grp[j]], classes[grp[i]], classes[grp[j]], class_aware, thr)
            # step B: greedy over the alive word with the masks
            while aw.any() and len(kept) < max_det:
                i = int(np.argmax(aw))
                kept.append(int(grp[i]))
                aw[i] = False
                aw &= ~pair[i]
    return np.asarray(kept, np.int64)


def random_case(rng, n, n_classes=5, dense=True):
    cx, cy = rng.random(n), rng.random(n)
    w, h = (0.05 + 0.3 * rng.random(n), 0.05 + 0.3 * rng.random(n)) if dense else (0.01 + 0.03 * rng.random(n), 0.01 + 0.03 * rng.random(n))
    boxes = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], -1).astype(F32)
    scores = rng.random(n).astype(F32)
    scores[rng.integers(0, n, n // 6)] = scores[rng.integers(0, n, n // 6)]       # ties
    classes = rng.integers(0, n_classes, n)
    return boxes, scores, classes


@pytest.mark.parametrize("chunk,group", [(16, 4), (32, 8), (64, 32), (1024, 32)])
def test_walk_equals_the_agnostic_reference_loop(chunk, group):
    rng = np.random.default_rng(chunk + group)
    for n, dense, cap in [(1, True, 5), (7, True, 100), (150, True, 100), (150, True, 9), (260, False, 40), (200, True, 1)]:
        boxes, scores, classes = random_case(rng, n, dense=dense)
        if n > 20:
            boxes[rng.integers(0, n, 3)] = np.nan                  # a NaN IoU suppresses in this rule
        want = detect_ref.nms_agnostic(boxes, scores, 0.45, cap)
        got = walk(boxes, scores, classes, 0.45, cap, False, chunk, group)
        assert got.tolist() == want.tolist(), (n, dense, cap)


@pytest.mark.parametrize("chunk,group", [(16, 4), (64, 32)])
def test_walk_equals_the_class_aware_reference_loop(chunk, group):
    rng = np.random.default_rng(7 * chunk + group)
    for n, dense, cap in [(5, True, 100), (180, True, 1000), (180, True, 12), (240, False, 50)]:
        boxes, scores, classes = random_case(rng, n, n_classes=3, dense=dense)
        if n > 20:
            boxes[rng.integers(0, n, 3)] = np.nan                  # a NaN IoU does NOT suppress in this rule
        want = detect_ref.nms_class_aware(boxes, scores, classes, 0.5, cap, boxes_are_corners=True)
        got = walk(boxes, scores, classes, 0.5, cap, True, chunk, group)
        assert got.tolist() == want.tolist(), (n, dense, cap)


def test_walk_with_score_threshold_and_no_survivors():
    rng = np.random.default_rng(2)
    boxes, scores, classes = random_case(rng, 120)
    scores[5] = np.nan
    passing = scores > F32(0.6)
    want = np.nonzero(passing)[0][detect_ref.nms_agnostic(boxes[passing], scores[passing], 0.45, 20)]
    got = walk(boxes, scores, classes, 0.45, 20, False, 16, 4, score_thr=0.6)
    assert got.tolist() == want.tolist()
    assert walk(boxes, scores, classes, 0.45, 20, False, 16, 4, score_thr=2.0).tolist() == []
