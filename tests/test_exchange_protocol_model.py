"""CPU model of the shared-mesh exchange inside k_camera_backward_shared_exchange (csrc/nr_camera.cu): every rank
pushes (value, epoch) words into the slots [epoch parity][its rank] of every peer's receive buffer, then polls its
own buffer until the words of every peer carry this step's epoch, and only then moves on to the next step.  No
flags, no barriers: what keeps a word from being overwritten before it was read is that a rank cannot get more than
one step ahead of a peer.  The model runs the protocol under adversarial interleavings (any rank may advance at any
time, pushes to different peers are separate events) and checks that every read returns the value the peer pushed
in THAT step.  (The kernel itself is checked on 2 and 8 GPUs by tools/check_shared_mesh_nccl.py.)"""
import random

import pytest


def run(world, steps, seed, parity_slots=2):
    rng = random.Random(seed)
    # recv[r][parity][src] = (value, epoch)
    recv = [[[(None, 0) for _ in range(world)] for _ in range(parity_slots)] for _ in range(world)]
    # program counter of a rank: (step e, phase, index): phase 0 = pushes still to do, phase 1 = peers still to read
    state = [dict(e=1, todo_push=[p for p in range(world) if p != r], todo_read=[p for p in range(world) if p != r], got={})
             for r in range(world)]
    value = lambda r, e: (r, e)                          # what rank r contributes in step e
    done = [0] * world
    while min(done) < steps:
        r = rng.randrange(world)
        st = state[r]
        if done[r] >= steps:
            continue
        e = st["e"]
        if st["todo_push"]:
            p = st["todo_push"].pop(rng.randrange(len(st["todo_push"])))
            recv[p][e % parity_slots][r] = (value(r, e), e)          # one atomic 8-byte store
            continue
        if st["todo_read"]:
            p = rng.choice(st["todo_read"])
            val, ep = recv[r][e % parity_slots][p]                   # one atomic 8-byte load
            assert ep <= e, "rank %d, step %d: the word of rank %d was overwritten by step %d before it was read" % (r, e, p, ep)
            if ep != e:
                continue                                             # not there yet (a stale word of step e - 2): poll again
            assert val == value(p, e), "rank %d read %r in step %d from rank %d" % (r, val, e, p)
            st["got"][p] = val
            st["todo_read"].remove(p)
            continue
        # the sum, in rank order, over exactly this step's contributions
        assert sorted(st["got"]) == [p for p in range(world) if p != r]
        done[r] += 1
        st.update(e=e + 1, todo_push=[p for p in range(world) if p != r], todo_read=[p for p in range(world) if p != r], got={})
    return True


@pytest.mark.parametrize("world", [2, 3, 8])
def test_every_read_returns_this_steps_value(world):
    for seed in range(20):
        assert run(world, steps=12, seed=seed)


def test_one_slot_per_source_would_not_be_enough():
    """Without the epoch parity a fast rank overwrites a word its peer has not read yet: the model must catch that
    (it is the reason the receive buffers hold two slots per source)."""
    failures = 0
    for seed in range(40):
        try:
            run(3, steps=12, seed=seed, parity_slots=1)
        except AssertionError:
            failures += 1
    assert failures > 0
