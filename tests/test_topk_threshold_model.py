"""CPU model of the K4 row scan (tricolo_b200/csrc/topk.cu: topk_rank_kernel): 32 lanes, 512-column chunks, per-lane
sorted lists with strictly-greater insertion, and the EXACT warp-wide k-th value as entry threshold, refreshed on the
kernel's schedule (after 1, 2, 4, 8 chunks, then every 8, and once before the tails).  The claim under test is the one the
kernel relies on: a value <= the k-th largest value of all EARLIER columns can never enter the row's top-k under the
stated order (similarity descending, gallery index ascending), so skipping it loses nothing - including on ties.
The GPU tests check the kernel itself bit for bit; this checks the rule on the CPU, with heavy ties."""
import numpy as np
import pytest

NEG = -np.finfo(np.float32).max
IMAX = np.iinfo(np.int32).max


def _push(v, i, x, idx):
    """TopList::push: strictly-greater insertion into a descending list (ties never displace an earlier entry)."""
    k = len(v)
    if not x > v[k - 1]:
        return
    v[k - 1], i[k - 1] = x, idx
    for t in range(k - 1, 0, -1):
        if v[t] > v[t - 1]:
            v[t], v[t - 1] = v[t - 1], v[t]
            i[t], i[t - 1] = i[t - 1], i[t]


def _refresh(lists, k):
    """k rounds of a warp-wide maximum over per-lane cursors: the exact k-th largest value seen so far."""
    cur = [0] * 32
    t = NEG
    for _ in range(k):
        cand = [lists[l][0][cur[l]] if cur[l] < len(lists[l][0]) else NEG for l in range(32)]
        t = max(cand)
        cur[cand.index(t)] += 1  # lowest lane holding the maximum advances
    return t


def model_topk(row, k, K=5):
    n = len(row)
    lists = [([NEG] * K, [IMAX] * K) for _ in range(32)]
    thr = NEG
    n_vec = n // 4
    n_full = n_vec // 128
    skipped = 0
    for ch in range(n_full):
        if ch != 0 and ((ch & (ch - 1)) == 0 or (ch & 7) == 0):
            thr = _refresh(lists, k)
        base4 = ch * 128
        for lane in range(32):
            cols = [(base4 + lane + j * 32) * 4 + u for j in range(4) for u in range(4)]
            vals = [row[c] for c in cols]
            if max(vals) > thr:
                for x, c in zip(vals, cols):
                    if x > thr:
                        _push(*lists[lane], x, c)
                    else:
                        skipped += 1
            else:
                skipped += 16
    if n_full > 0:
        thr = _refresh(lists, k)
    g4 = n_full * 128
    while g4 < n_vec:  # float4 tail, 32 groups at a time
        for lane in range(32):
            if g4 + lane < n_vec:
                for u in range(4):
                    c = (g4 + lane) * 4 + u
                    if row[c] > thr:
                        _push(*lists[lane], row[c], c)
        g4 += 32
    for lane in range(32):  # scalar tail
        c = n_vec * 4 + lane
        if c < n:
            _push(*lists[lane], row[c], c)
    # k rounds of warp arg-best over the lane heads under (value desc, index asc)
    out = []
    for _ in range(k):
        best = None
        for lane in range(32):
            v, i = lists[lane]
            if best is None or v[0] > best[0] or (v[0] == best[0] and i[0] < best[1]):
                best = (v[0], i[0], lane)
        out.append((best[0], best[1] if best[1] != IMAX else -1))
        v, i = lists[best[2]]
        if best[1] != IMAX:
            v.pop(0); i.pop(0); v.append(NEG); i.append(IMAX)
    return out, skipped


def exact_topk(row, k):
    order = sorted(range(len(row)), key=lambda c: (-row[c], c))
    return [(row[c], c) for c in order[:k]]


@pytest.mark.parametrize("n,k,kind", [(2048, 5, "gauss"), (5000, 5, "ties"), (4613, 3, "ties"), (25000, 5, "gauss"),
                                      (9000, 1, "ties"), (700, 5, "gauss"), (12288, 5, "ascending"), (6144, 5, "constant")])
def test_threshold_rule_keeps_exact_topk(n, k, kind):
    rng = np.random.default_rng(n + k)
    if kind == "gauss":
        row = rng.standard_normal(n).astype(np.float32)
    elif kind == "ties":
        row = rng.integers(0, 7, n).astype(np.float32)  # a handful of distinct values: ties everywhere
    elif kind == "ascending":
        row = np.arange(n, dtype=np.float32)  # every value beats everything before it: the threshold never helps
    else:
        row = np.full(n, 0.25, dtype=np.float32)
    got, skipped = model_topk([float(x) for x in row], k)
    want = exact_topk([float(x) for x in row], k)
    assert got == want
    if kind in ("gauss", "ties", "constant") and n >= 4096:
        assert skipped > 0.6 * (n // 512) * 512  # and the rule removes most of the list work (the first chunks cannot skip)
