#!/usr/bin/env python
"""Model of the HBM kernel's per-episode update (th_rl_b200/csrc/thrl_scan_hbm.cuh) checked against the plain sequential
form of QTable.train_net (th_rl/agents.py:59-78: stale snapshot, live row max, writes in batch order).

The kernel never walks the batch with a load -> max -> store chain through HBM.  It
  1. snapshots old[j] = Q[s_j, k_j] for every transition j (before anything is written),
  2. TAGS every cell the batch will write: the cell temporarily holds a marker carrying the smallest transition index that
     writes it (atomic max of 1023 - j over a NaN-boxed payload) -- the table itself says which cells are being rewritten,
  3. gathers the rows of all states of the episode with bulk copies (all loads independent),
  4. walks the states in order entirely on chip: the gathered row with its tagged cells replaced by their current values
     cur[canonical index] IS the live row (agents.py:71), so next_max is one reduction; the new value goes to cur[],
  5. writes cur[] of the canonical transitions back over the tags, and refreshes the greedy-action cache of every touched
     row exactly: (max, first argmax) over the row's untagged cells (immune to this batch) merged with the final values of
     its tagged cells.
This script replays that on random batches with many repeated cells / rows and compares tables and greedy actions.
"""
import numpy as np

TAG = 1 << 40  # any value no table holds; the payload is added to it


def reference_update(Q, s, k, r, alpha, gamma):
    Q = Q.copy()
    old = Q[s[:-1], k].copy()                      # agents.py:67 (snapshot)
    for j in range(len(k)):                        # agents.py:68-76
        nm = Q[s[j + 1]].max()                     # live
        Q[s[j], k[j]] = (1 - alpha) * old[j] + alpha * (r[j] + gamma * nm)
    return Q


def kernel_update(Q, s, k, r, alpha, gamma):
    Q = Q.copy()
    L = len(k)
    cur = Q[s[:-1], k].copy()                      # 1. snapshot, one slot per transition
    for j in range(L):                             # 2. tag: smallest j wins
        c = Q[s[j], k[j]]
        Q[s[j], k[j]] = TAG + max(c - TAG if c >= TAG else -1, 1023 - j)
    rows = [Q[s[t]].copy() for t in range(L + 1)]  # 3. gather (tags included)
    canon = np.zeros(L, int)
    base = []
    for t in range(L + 1):                         # 4. on-chip walk
        row = rows[t]
        tagged = row >= TAG
        idx = (1023 - (row[tagged] - TAG)).astype(int)
        un = np.where(tagged, -np.inf, row)
        bm = un.max()
        ba = int(np.argmax(un)) if np.isfinite(bm) else None
        base.append((bm, ba))
        live = max([bm] + [cur[c] for c in idx])
        if t >= 1:
            j = t - 1
            nv = (1 - alpha) * cur[j] + alpha * (r[j] + gamma * live)   # cur[j] is still the snapshot (canonical = first writer)
            cur[canon[j]] = nv
        if t < L:
            canon[t] = int(1023 - (row[k[t]] - TAG))
    for j in range(L):                             # 5a. write back over the tags
        if canon[j] == j:
            Q[s[j], k[j]] = cur[j]
    greedy = {}
    for t in range(L + 1):                         # 5b. exact greedy refresh of every row of the episode
        bm, ba = base[t]
        cand = [(bm, ba)] if ba is not None else []
        cand += [(cur[j], k[j]) for j in range(L) if canon[j] == j and s[j] == s[t]]
        best = max(v for v, _ in cand)
        greedy[s[t]] = min(c for v, c in cand if v == best)
    return Q, greedy


def main():
    rng = np.random.default_rng(0)
    for case in range(3000):
        S, A = int(rng.integers(2, 12)), int(rng.integers(2, 9))
        L = int(rng.integers(1, 60))
        Q = rng.normal(10, 1, (S, A))
        if case % 3 == 0:
            Q = np.round(Q)                        # many exact ties
        s = rng.integers(0, S if case % 2 else min(S, 3), L + 1)
        k = rng.integers(0, A, L)
        r = rng.normal(1, 1, L).round(1)
        alpha, gamma = float(rng.choice([0.05, 0.5, 1.0])), float(rng.choice([0.0, 0.9, 0.99]))
        want = reference_update(Q, s, k, r, alpha, gamma)
        got, greedy = kernel_update(Q, s, k, r, alpha, gamma)
        assert np.array_equal(want, got), case
        for row, g in greedy.items():
            assert g == int(np.argmax(want[row])), (case, row)
    print("model == sequential train_net on 3000 random batches (tables bit-equal, greedy cache exact)")


if __name__ == "__main__":
    main()
