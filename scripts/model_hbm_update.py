#!/usr/bin/env python
"""Model of the HBM kernel's per-episode update (th_rl_b200/csrc/thrl_scan_hbm.cuh, steps U2-U5) checked against the plain
sequential form of QTable.train_net (th_rl/agents.py:59-78: stale snapshot, live row max, writes in batch order).

The kernel never walks the batch with a load -> max -> store chain through HBM.  Per episode and agent it
  U2  groups the episode's states by table row (ascending chain per row, the first state is the row's representative),
  U3  gathers every DISTINCT row once (bulk copies, all independent); per gathered row the transitions of that row are
      walked in order: the first writer of a cell takes the staged value as the stale snapshot of every writer of the cell
      and TAGS the staged cell; then (max, first argmax) over the row's untagged cells is taken -- no write of this batch
      can change it,
  U4  walks the batch in order on chip: live row max = untagged max of the next state's row merged with the current values
      of that row's rewritten cells; the new value goes to the slot of the cell's first writer,
  U5  stores the final value of every rewritten cell and refreshes the greedy action of every touched row from the untagged
      (max, argmax) and the final values of the rewritten cells.
This script replays that on random batches with many repeated cells / rows and compares tables and greedy actions.
"""
import numpy as np

NONE = 0xFF


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
    # U2: chains through the states of one row, built in descending order (the head ends up being the first state)
    head, nexts = {}, [NONE] * (L + 1)
    for t in range(L, -1, -1):
        nexts[t] = head.get(s[t], NONE)
        head[s[t]] = t
    rs = [head[s[t]] for t in range(L + 1)]
    reps = [t for t in range(L + 1) if rs[t] == t]
    cur = np.zeros(L)
    canon = [0] * L
    nextc = [NONE] * L
    headc, bm, ba = {}, {}, {}
    for rep in reps:                               # U3: one gathered copy per distinct row
        staged = Q[s[rep]].copy()
        tag = {}
        h = NONE
        t = rep
        while t != NONE and t < L:
            if k[t] in tag:
                f = tag[k[t]]
                cur[t], canon[t] = cur[f], f
            else:
                cur[t], canon[t] = staged[k[t]], t
                tag[k[t]] = t
                nextc[t], h = h, t
            t = nexts[t]
        headc[rep] = h
        un = np.array([-np.inf if c in tag else staged[c] for c in range(len(staged))])
        free = [c for c in range(len(staged)) if c not in tag]
        bm[rep] = un.max()
        ba[rep] = (int(np.argmax(un)) if np.isfinite(bm[rep]) or free else None) if free else None

    def cells(rep):
        c = headc[rep]
        while c != NONE:
            yield c
            c = nextc[c]

    for j in range(L):                             # U4
        rep = rs[j + 1]
        m = max([bm[rep]] + [cur[c] for c in cells(rep)])
        nv = (1 - alpha) * cur[j] + alpha * (r[j] + gamma * m)
        cur[canon[j]] = nv
    greedy = {}
    for t in range(L + 1):                         # U5
        if t < L and canon[t] == t:
            Q[s[t], k[t]] = cur[t]
        if rs[t] == t:
            best, bidx = bm[t], (ba[t] if ba[t] is not None else 1 << 30)
            for c in cells(t):
                if cur[c] > best or (cur[c] == best and k[c] < bidx):
                    best, bidx = cur[c], k[c]
            greedy[s[t]] = bidx
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
