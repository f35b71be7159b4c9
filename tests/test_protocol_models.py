"""Model checks (pure Python, no GPU) of the two flag protocols the partitioned step of round 2 relies on, under
randomised interleavings of the ranks' atomic actions:

* the dt exchange of k_adv_p2p (ftb200_kernels.cuh, P2PHeader::dtslot): ONE store per rank and step, the value being its
  own arrival flag, four rotating slots, the reader re-arming slot (s + 2) % 4 after consuming slot s % 4;
* the fused shared-node exchange of elem_p2p_epilogue: every boundary element counts itself in at its shared nodes, the
  last arrival packs the node, the thread that packs the last node raises the flag -- exactly once per step, after every
  node has been packed, whatever the order of arrival.

These are models of the protocols, not of the CUDA code: they pin the ordering argument written next to the code."""
import random

import pytest

EMPTY = None


class Rank:
    """One rank of the dt exchange as a state machine; `slots[k][p]` is this rank's window."""

    def __init__(self, r, P, steps, rng):
        self.r, self.P, self.steps = r, P, steps
        self.slots = [[EMPTY] * P for _ in range(4)]
        self.seq = 0
        self.pc = 0          # 0 publish, 1 wait/read, 2 re-arm, then the rest of the step
        self.todo = []       # peers still to be written in the publish phase
        self.seen = []       # per step: tuple of the P values read
        self.rng = rng

    def mydt(self, seq):
        return (self.r, seq)  # unique per (rank, step): a stale or lost value is detected

    def done(self):
        return self.seq >= self.steps

    def act(self, ranks):
        """One atomic action; returns False if the rank is blocked (waiting)."""
        if self.pc == 0:
            if not self.todo:
                self.todo = list(range(self.P))
                self.rng.shuffle(self.todo)
            p = self.todo.pop()
            s = ranks[p].slots[self.seq & 3]
            assert s[self.r] is EMPTY, "peer's slot still holds an unread value: a dt would be lost"
            s[self.r] = self.mydt(self.seq)
            if not self.todo:
                self.pc = 1
            return True
        if self.pc == 1:
            cur = self.slots[self.seq & 3]
            if any(v is EMPTY for v in cur):
                return False
            self.seen.append(tuple(cur))
            self.pc = 2
            return True
        # re-arm the slot two steps ahead, then the step is over
        self.slots[(self.seq + 2) & 3] = [EMPTY] * self.P
        self.seq += 1
        self.pc = 0
        return True


@pytest.mark.parametrize("P", [1, 2, 3, 8])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_dt_slots_value_is_its_own_flag(P, seed):
    rng = random.Random(100 * P + seed)
    steps = 40
    ranks = [Rank(r, P, steps, rng) for r in range(P)]
    guard = 0
    while not all(k.done() for k in ranks):
        live = [k for k in ranks if not k.done()]
        k = rng.choice(live)
        k.act(ranks)
        guard += 1
        assert guard < 10 ** 6, "deadlock"
    for k in ranks:
        assert len(k.seen) == steps
        for s, vals in enumerate(k.seen):
            assert vals == tuple((p, s) for p in range(P))  # every peer's value of exactly this step


def test_dt_slots_model_detects_a_missing_rearm():
    """Non-vacuity: the same protocol WITHOUT the re-arm leaves the value of step s in the slot that step s + 4 polls, so a
    reader can take a stale dt for the new one (or a writer finds the slot occupied) -- the model must notice."""
    class NoRearm(Rank):
        def act(self, ranks):
            if self.pc == 2:
                self.seq += 1
                self.pc = 0
                return True
            return super().act(ranks)
    hit = False
    for seed in range(20):
        rng = random.Random(seed)
        ranks = [NoRearm(r, 2, 12, rng) for r in range(2)]
        try:
            for _ in range(100000):
                live = [k for k in ranks if not k.done()]
                if not live:
                    break
                rng.choice(live).act(ranks)
            for k in ranks:
                for st, vals in enumerate(k.seen):
                    assert vals == tuple((p, st) for p in range(2))
        except AssertionError:
            hit = True
            break
    assert hit


@pytest.mark.parametrize("seed", range(5))
def test_last_arrival_packs_each_shared_node_once_and_flags_once(seed):
    """elem_p2p_epilogue: elements arrive at their shared nodes in arbitrary order (threads of many blocks); per step every
    shared node is packed exactly once, by its last contributor, after all of its contributions are visible, and the flag
    is raised exactly once, after the last pack.  Counters return to zero for the next step."""
    rng = random.Random(seed)
    n_shared, n_elem = 37, 60
    # element -> the shared nodes it touches (1..4 of them), every node touched by at least one element
    elems = [rng.sample(range(n_shared), rng.randint(1, 4)) for _ in range(n_elem)]
    for h in range(n_shared):
        if not any(h in e for e in elems):
            elems[rng.randrange(n_elem)].append(h)
    deg = [sum(e.count(h) for e in elems) for h in range(n_shared)]
    ctr = [0] * n_shared
    packed = [0]
    for step in range(3):
        stored = set()          # elements whose force stores are visible
        pack_log, flags = [], []
        # each element: store forces, then count in at all nodes (phase A), then pack its nodes (phase B), then add to `packed`
        actions = []
        for e in range(n_elem):
            actions.append([("store", e)] + [("arrive", e, h) for h in elems[e]] + [("finish", e)])
        mine = {e: [] for e in range(n_elem)}
        while any(actions):
            q = rng.choice([a for a in actions if a])
            act = q.pop(0)
            if act[0] == "store":
                stored.add(act[1])
            elif act[0] == "arrive":
                _, e, h = act
                ctr[h] += 1
                if ctr[h] == deg[h]:
                    ctr[h] = 0
                    mine[e].append(h)
            else:
                e = act[1]
                for h in mine[e]:
                    assert all(o in stored for o in range(n_elem) if h in elems[o]), "packed before a contribution was visible"
                    pack_log.append(h)
                if mine[e]:
                    packed[0] += len(mine[e])
                    if packed[0] == n_shared:
                        packed[0] = 0
                        flags.append(len(pack_log))
        assert sorted(pack_log) == list(range(n_shared))
        assert flags == [n_shared]
        assert ctr == [0] * n_shared and packed == [0]
