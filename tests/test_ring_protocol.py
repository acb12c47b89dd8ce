"""Model check of the ring kernel's producer / consumer protocol (sd_blkr.h, `Ring protocol` in its header):
entries e = 0, 1, 2, ... in slot e % NB, full[] / empty[] mbarriers waited on with phase parity, tile headers in
hdr[t % NB] written after the wait on empty[] of the tile's first entry, end-of-list sentinel.

The kernel itself needs a GPU; what can be checked without one is the protocol.  This file restates the two loops
of sd_blkr_apply_kernel as coroutines over a faithful mbarrier model (pending-arrival count, transaction bytes,
phase bit, try_wait.parity semantics) with TMA completions delivered at random later times, runs them under many
random interleavings and asserts:
  * no deadlock, every tile processed exactly once by every consumer warp;
  * a consumer never reads a ring slot or a header that holds anything but the entry / tile it expects;
  * the producer never overwrites a slot or a header some consumer warp has not released yet.
A deliberately broken variant (header written BEFORE the wait) must be caught, so the checker has teeth."""
import random

import pytest

NB = 4


class MBar:
    def __init__(self, count):
        self.count = count
        self.pending = count
        self.tx = 0
        self.phase = 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self):
        assert self.pending > 0, "more arrivals than the barrier expects in one phase"
        self.pending -= 1
        self._check()

    def arrive_expect_tx(self, nbytes):
        assert self.pending > 0
        self.tx += nbytes
        self.pending -= 1
        self._check()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._check()

    def try_wait(self, parity):        # true once the phase of that parity has completed
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, tiles, ncons, rng, hdr_before_wait=False):
        self.tiles = tiles              # ntot per tile of this CTA
        self.ncons = ncons
        self.rng = rng
        self.full = [MBar(1) for _ in range(NB)]
        self.empty = [MBar(ncons) for _ in range(NB)]
        self.slot = [None] * NB         # (tile, n) the slot holds
        self.hdr = [None] * NB          # tile number (or "end") the header holds
        self.reading_slot = [set() for _ in range(NB)]
        self.reading_hdr = [set() for _ in range(NB)]
        self.tma = []                   # in-flight copies: (slot, tag, bytes)
        self.done = [[] for _ in range(ncons)]
        self.hdr_before_wait = hdr_before_wait

    def write_hdr(self, t, value):
        assert not self.reading_hdr[t % NB], f"header {t % NB} overwritten while warps {self.reading_hdr[t % NB]} use it"
        self.hdr[t % NB] = value

    def producer(self):
        e = 0
        t = 0
        while True:
            end = t >= len(self.tiles)
            if self.hdr_before_wait:
                self.write_hdr(t, "end" if end else t)
            while not self.empty[e % NB].try_wait(((e // NB) & 1) ^ 1):
                yield
            if not self.hdr_before_wait:
                self.write_hdr(t, "end" if end else t)
            if end:
                self.full[e % NB].arrive()
                return
            ntot = self.tiles[t]
            for n in range(ntot + 1):
                s = e % NB
                if n > 0:
                    while not self.empty[s].try_wait(((e // NB) & 1) ^ 1):
                        yield
                assert not self.reading_slot[s], f"slot {s} refilled while warps {self.reading_slot[s]} read it"
                nbytes = 7 + (n % 3)
                self.full[s].arrive_expect_tx(nbytes)
                for _ in range(nbytes):                     # several bulk copies per entry, each lands on its own
                    self.tma.append((s, (t, n), 1))
                e += 1
                yield
            t += 1

    def consumer(self, w):
        e = 0
        t = 0
        while True:
            while not self.full[e % NB].try_wait((e // NB) & 1):
                yield
            h = t % NB
            if self.hdr[h] == "end":
                return
            assert self.hdr[h] == t, f"warp {w} expected header of tile {t}, found {self.hdr[h]}"
            self.reading_hdr[h].add(w)
            ntot = self.tiles[t]
            for n in range(ntot + 1):
                s = e % NB
                if n > 0:
                    while not self.full[s].try_wait((e // NB) & 1):
                        yield
                assert self.slot[s] == (t, n), f"warp {w} expected entry {(t, n)} in slot {s}, found {self.slot[s]}"
                self.reading_slot[s].add(w)
                yield                                       # the LDS / DFMA work on the slot
                assert self.slot[s] == (t, n) and self.hdr[h] == t
                self.reading_slot[s].discard(w)
                if n == ntot:
                    self.reading_hdr[h].discard(w)          # last use of the header precedes the arrive on the own entry
                    self.done[w].append(t)
                self.empty[s].arrive()
                e += 1
                yield
            t += 1

    def run(self):
        threads = {"p": self.producer()}
        threads.update({w: self.consumer(w) for w in range(self.ncons)})
        idle = 0
        while threads:
            choices = list(threads) + (["tma"] if self.tma else [])
            pick = self.rng.choice(choices)
            if pick == "tma":
                s, tag, nbytes = self.tma.pop(self.rng.randrange(len(self.tma)))
                assert not self.reading_slot[s]
                self.slot[s] = tag
                self.full[s].complete_tx(nbytes)
                idle = 0
                continue
            before = self._state()
            try:
                next(threads[pick])
            except StopIteration:
                del threads[pick]
            idle = idle + 1 if self._state() == before and not self.tma else 0
            assert idle < 20000, "deadlock"
        assert not self.tma
        for w in range(self.ncons):
            assert self.done[w] == list(range(len(self.tiles)))

    def _state(self):
        return (tuple(b.phase for b in self.full), tuple(b.phase for b in self.empty),
                tuple(b.pending for b in self.empty), tuple(len(d) for d in self.done))


@pytest.mark.parametrize("seed", range(40))
def test_ring_protocol_random_interleavings(seed):
    rng = random.Random(seed)
    ntiles = rng.randrange(0, 14)
    # ntot = 0 (tile with no active prefix bond), 1, ... up to 18 (17 prefix bonds + crossing)
    tiles = [rng.choice([0, 0, 1, 1, 2, 3, 5, 9, 18]) for _ in range(ntiles)]
    Sim(tiles, rng.choice([1, 2, 3, 15]), rng).run()


def test_ring_protocol_all_single_entry_tiles():
    """Tiles with only their own entry are the tight case of the header-reuse argument (NB tiles per ring turn)."""
    for seed in range(20):
        Sim([0] * 23, 3, random.Random(1000 + seed)).run()


def test_checker_catches_a_header_written_before_the_wait():
    caught = 0
    for seed in range(30):
        try:
            Sim([0] * 12, 3, random.Random(seed), hdr_before_wait=True).run()
        except AssertionError:
            caught += 1
    assert caught > 0
