"""List-based restatement of ``mu::composePieceWisePol`` (reference neptune/src/utils.cpp:318-402) used to
cross-check the C oracle, plus random committed-trajectory records for the compose tests."""
import numpy as np

TP, REC, REC_PWP = 16, 256, 210


def rec_from(times, coeff):
    """times [n+1], coeff [3][n][4] -> record (256 doubles, header left zero)."""
    n = len(times) - 1
    r = np.zeros(REC)
    r[0] = n
    r[1:2 + n] = times
    c = r[1 + TP + 1:REC_PWP].reshape(3, TP, 4)
    c[:, :n] = coeff
    return r


def rec_to(r):
    n = int(r[0])
    return list(r[1:2 + n]), r[1 + TP + 1:REC_PWP].reshape(3, TP, 4)[:, :n].copy()


def compose_lists(t, p1, p2):
    """p1, p2 = (times list, coeff [3][n][4]); returns (times, list of [3][4] pieces) or ([], [])."""
    t1, c1 = list(p1[0]), p1[1]
    t2, c2 = list(p2[0]), p2[1]
    if t > t1[-1] and t < t2[0]:          # :320-324
        t2[0] = t
    if t1[-1] < t2[0]:                    # :326-330
        t2[0] = t1[-1]
    if t < t1[0]:                         # :332-336
        t1[0] = t
    if abs(t - t2[0]) < 1e-5:             # :338-341
        return t2, [c2[:, i] for i in range(len(t2) - 1)]
    if t1[-1] < t2[0] or t > t2[-1] or t < t1[0]:   # :343-354
        return [], []
    idx1 = [i for i in range(len(t1)) if t1[i] > t and t1[i] < t2[0]]   # :356-365
    idx2 = [i for i in range(len(t2)) if t2[i] > t]                     # :367-374
    times, pieces = [t], []
    for i in idx1:                        # :378-385
        times.append(t1[i])
        pieces.append(c1[:, i - 1])
    for i in idx2:                        # :387-399
        times.append(t2[i])
        pieces.append(c1[:, -1] if i == 0 else c2[:, i - 1])
    return times, pieces


def random_case(rng, kind):
    """One (t, prev record, now record) pair.  kind selects the branch of the reference function."""
    n1 = int(rng.integers(1, 9))
    n2 = int(rng.integers(1, 9))
    T = 0.5
    a = float(rng.uniform(0.0, 100.0))
    t1 = a + T * np.arange(n1 + 1)
    c1 = rng.normal(size=(3, n1, 4))
    c2 = rng.normal(size=(3, n2, 4))
    if kind == "mid":         # replan started somewhere inside prev, now starts later inside prev
        t = float(rng.uniform(t1[0], t1[-1] - 1e-3))
        s2 = float(rng.uniform(t + 2e-5, t1[-1]))
    elif kind == "same":      # now starts at t (within 1e-5)
        t = float(rng.uniform(t1[0], t1[-1]))
        s2 = t + float(rng.uniform(-9e-6, 9e-6))
    elif kind == "gap":       # t beyond the end of prev, now later still
        t = float(t1[-1] + rng.uniform(1e-3, 1.0))
        s2 = t + float(rng.uniform(1e-3, 1.0))
    elif kind == "late_now":  # now starts after prev ends
        t = float(rng.uniform(t1[0], t1[-1]))
        s2 = float(t1[-1] + rng.uniform(1e-3, 1.0))
    elif kind == "early":     # t before prev starts
        t = float(t1[0] - rng.uniform(1e-3, 1.0))
        s2 = float(rng.uniform(t1[0], t1[-1]))
    elif kind == "stale":     # t beyond the end of now
        s2 = float(rng.uniform(t1[0], t1[-1]))
        t = float(s2 + T * n2 + rng.uniform(1e-3, 1.0))
    elif kind == "knot":      # t and the start of now exactly on knots of prev
        k = int(rng.integers(0, n1 + 1))
        t = float(t1[k])
        s2 = float(t1[int(rng.integers(k, n1 + 1))])
    else:
        raise KeyError(kind)
    t2 = s2 + T * np.arange(n2 + 1)
    return t, rec_from(t1, c1), rec_from(t2, c2)


KINDS = ["mid", "same", "gap", "late_now", "early", "stale", "knot"]
