"""A model of the step kernel's weight ring (csrc/decoder_batch.cu: "Ring protocol"), run under random schedules on the CPU.

The kernel streams every CTA's weights through a ring of n slots guarded by mbarriers whose waits name only the PARITY of the use
they wait for.  Two tile groups consume alternate uses and skip each other's slots without waiting, which leaves two ways for a
waiter to be a whole ring cycle away from its barrier:

  (early)  the previous use of the slot belonged to the other group and is still in flight: the barrier has not flipped yet, so
           a wait for the NEXT parity passes at once, on stale bytes, and its release corrupts the empty barrier;
  (late)   a group that waits on slots it does not release can arrive after the slot was released and refilled: the parity has
           come round and the wait never returns.

The kernel's rule - only the owner waits, and before its wait it checks seen[slot] (the index of the last use some waiter
watched complete) - must survive every interleaving; the two broken variants must be caught by the same model.  This is test
infrastructure: it documents why the rule is what it is (both failures were first met on the GPU: DESIGN.md section 4.4)."""
import random

import pytest


class Barrier:
    """mbarrier as the kernel uses it: a phase bit that flips when the pending count reaches zero."""

    def __init__(self, count):
        self.count, self.pending, self.bit = count, count, 0

    def arrive(self):
        assert self.pending > 0, "arrival on a completed phase (count corrupted)"
        self.pending -= 1
        if self.pending == 0:
            self.bit ^= 1
            self.pending = self.count

    def try_wait(self, parity):              # "has the phase with this parity completed?"
        return self.bit != parity


def simulate(owners, n_slots, rule, rng, max_ticks=200000):
    """owners[u] in {0, 1}: the tile group that consumes use u.  rule: 'seen' (the kernel), 'none' (no seen check), 'both-wait' (the
    non-owner also waits on every slot).  Returns 'ok', 'early' (a wait passed before its data had arrived) or 'deadlock'."""
    full = [Barrier(1) for _ in range(n_slots)]
    empty = [Barrier(1) for _ in range(n_slots)]
    seen = [-1] * n_slots
    arrived = [False] * len(owners)          # ground truth: the data of use u is in the slot
    in_flight = []                           # uses whose copy has been issued but has not landed
    prod = 0                                 # next use the producer issues
    pos = [0, 0]                             # next use each group looks at
    consumed = [False] * len(owners)
    for _ in range(max_ticks):
        if prod >= len(owners) and all(p >= len(owners) for p in pos):
            return "ok" if all(consumed) else "deadlock"
        actor = rng.choice(("producer", "land", "g0", "g1"))
        if actor == "producer" and prod < len(owners):
            s, phase = prod % n_slots, (prod // n_slots) & 1
            if empty[s].try_wait(phase ^ 1):                         # the slot's previous use has been released
                in_flight.append(prod)
                prod += 1
        elif actor == "land" and in_flight:
            u = in_flight.pop(rng.randrange(len(in_flight)))         # copies land in any order
            arrived[u] = True
            full[u % n_slots].arrive()
        elif actor in ("g0", "g1"):
            g = int(actor[1])
            u = pos[g]
            if u >= len(owners):
                continue
            s, phase = u % n_slots, (u // n_slots) & 1
            if owners[u] != g:
                if rule == "both-wait" and not full[s].try_wait(phase):
                    continue                                         # the non-owner waits too (and never releases)
                pos[g] += 1
                continue
            if rule == "seen" and seen[s] < u - n_slots:
                continue                                             # db_seen_wait
            if not full[s].try_wait(phase):
                continue
            if not arrived[u]:
                return "early"
            seen[s] = u
            consumed[u] = True
            empty[s].arrive()
            pos[g] += 1
    return "deadlock"


def _schedule_mlp2_tiles(n_tiles=6, slots_per_tile=4):
    """MLP2 on a lane: tiles of K = 4d span four slots each and alternate between the groups (the large-v3 hang)."""
    return [t & 1 for t in range(n_tiles) for _ in range(slots_per_tile)]


def _schedule_single_slot_tiles(n=24):
    return [u & 1 for u in range(n)]


@pytest.mark.parametrize("n_slots", [2, 3, 4])
@pytest.mark.parametrize("owners", [_schedule_mlp2_tiles(), _schedule_single_slot_tiles(), [0] * 9 + _schedule_mlp2_tiles(4) + [1, 0, 1]])
def test_seen_rule_survives_every_interleaving(n_slots, owners):
    for seed in range(300):
        assert simulate(owners, n_slots, "seen", random.Random(seed)) == "ok", (n_slots, seed)


def test_without_the_seen_check_a_wait_passes_on_stale_bytes():
    owners = _schedule_mlp2_tiles()
    outcomes = {simulate(owners, 4, "none", random.Random(seed)) for seed in range(400)}
    assert "early" in outcomes, outcomes


def test_a_waiter_that_does_not_release_can_wait_forever():
    owners = _schedule_single_slot_tiles()
    outcomes = {simulate(owners, 2, "both-wait", random.Random(seed), max_ticks=20000) for seed in range(400)}
    assert "deadlock" in outcomes, outcomes


# ---- the double-buffered stage descriptor ---------------------------------------------------------------------------------
def simulate_descriptor(has_barrier, rng, max_ticks=20000):
    """Thread 0 (in group 0) writes desc[it & 1] at the top of stage `it`; every group reads desc[it & 1] when it stores the outputs
    of its last tile of stage `it`.  has_barrier[it]: the stage passes a CTA-wide barrier (after the descriptor write).  Returns
    'ok' or 'stale' (a group stored stage it's outputs through another stage's descriptor)."""
    n = len(has_barrier)
    desc = [None, None]
    # per group: (stage, point) with point 0 = stage top, 1 = at the barrier, 2 = storing outputs
    st = [[0, 0], [0, 0]]
    for _ in range(max_ticks):
        if all(s[0] >= n for s in st):
            return "ok"
        g = rng.randrange(2)
        it, point = st[g]
        if it >= n:
            continue
        if point == 0:
            if g == 0:
                desc[it & 1] = it
            st[g][1] = 1
        elif point == 1:
            other = st[1 - g]
            if has_barrier[it] and not (other[0] > it or (other[0] == it and other[1] >= 1)):
                continue                                             # wait at the barrier for the other group
            st[g][1] = 2
        else:
            if desc[it & 1] != it:
                return "stale"
            st[g] = [it + 1, 0]
    return "ok"


def test_one_barrier_per_stage_keeps_the_descriptor_valid():
    for seed in range(500):
        assert simulate_descriptor([True] * 12, random.Random(seed)) == "ok", seed


def test_a_stage_without_a_barrier_lets_the_descriptor_be_overwritten():
    """The prompt launches before the fix: a CTA without a self-attention unit had no barrier in stage 1."""
    has = [True] * 12
    has[1] = False
    outcomes = {simulate_descriptor(has, random.Random(seed)) for seed in range(500)}
    assert "stale" in outcomes, outcomes
