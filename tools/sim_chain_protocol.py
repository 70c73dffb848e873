"""Randomised model check of the synchronisation protocol of acezero_b200/csrc/head_chain.cu (no GPU needed).

The kernel's correctness rests on a handful of mbarriers shared by four kinds of agents in each of the two CTAs of a
cluster (TMA producer, UMMA issuer, two epilogue groups) plus asynchronous engines (TMA loads / stores, DSMEM bulk
copies, the tensor core). This script replays exactly the wait / arrive / expect_tx sequence of the kernel under random
interleavings and random completion times of every asynchronous operation and asserts, at every step,
  * an mbarrier wait never passes on a stale or aliased phase (the phase that completed is the one it was meant for),
  * the UMMA of step s reads box j only while the box holds k-block j of the input of step s,
  * nobody writes a shared-memory box that an in-flight reader (UMMA, DSMEM copy, TMA store) still reads,
  * a TMEM buffer is not overwritten before the epilogue that reads it has finished, and is read only when complete,
  * the run ends (no deadlock) with every step's tile stored.
Run: python tools/sim_chain_protocol.py [runs] [n_steps]
"""
import random
import sys

KB, BST = 8, 3


class MBar:
    def __init__(self, name, count=1):
        self.name, self.count, self.pending, self.tx, self.phase = name, count, count, 0, 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self):
        assert self.pending > 0, f"{self.name}: arrive on a phase with no pending arrivals"
        self.pending -= 1
        self._check()

    def arrive_expect_tx(self, n):
        self.tx += n
        self.arrive()

    def complete_tx(self, n):
        self.tx -= n
        self._check()

    def passed(self, parity):
        return (self.phase & 1) != parity


class CTA:
    def __init__(self, rank):
        self.rank = rank
        self.a_ready = [MBar(f"c{rank}.a_ready{j}") for j in range(KB)]
        self.b_full = [MBar(f"c{rank}.b_full{i}") for i in range(BST)]
        self.b_empty = [MBar(f"c{rank}.b_empty{i}") for i in range(BST)]
        self.tmem_full = [MBar(f"c{rank}.tmem_full{i}") for i in range(2)]
        self.peer_free = MBar(f"c{rank}.peer_free")
        self.a_half = [MBar(f"c{rank}.a_half{b}") for b in range(4)]
        self.A_h0 = [None] * KB       # V3: tag of the FIRST half (32 columns) of the box
        self.A = [None] * KB          # tag: step whose INPUT k-block j the box currently holds
        self.A_readers = [0] * KB     # in-flight asynchronous readers of the box
        self.A_writers = [0] * KB
        self.B = [None] * BST         # (step, i)
        self.tmem = [None, None]      # ("acc", s, complete?) per buffer
        self.tmem_reading = [0, 0]
        self.mma_queue = []           # issued, not yet retired UMMA batches (retire in order)
        self.store_groups = []        # per issuer (group g): list of [reads_done?]
        self.stored = set()
        self.drained = {}             # step -> boxes whose TMEM read has finished


V3 = False  # set by main(): arrival-order consumption + own boxes published in two halves (head_chain.cu V3)


def order(i, rank):
    if V3:
        b = ((i >> 2) << 1) | (i & 1)
        return ((rank ^ 1) if (i & 2) else rank) * 4 + b
    return rank * 4 + i if i < 4 else (rank ^ 1) * 4 + (i - 4)


def is_peer_slot(i):
    return bool(i & 2) if V3 else i >= 4


class Sim:
    def __init__(self, n_steps, seed):
        self.n = n_steps
        self.rng = random.Random(seed)
        self.p_async = (0.01, 0.05, 0.35, 0.7)[seed % 4]  # how eagerly asynchronous engines complete
        self.c = [CTA(0), CTA(1)]
        self.pending_async = []  # callables that may fire at any later time
        self.agents = []
        for r in (0, 1):
            self.agents += [self.producer(r), self.mma(r), self.epi(r, 0), self.epi(r, 1)]
        self.bar3_gen = [[0, 0], [0, 0]]

    # ---- helper: wait for a specific absolute phase of a barrier ----
    def wait(self, bar, parity, expect_phase):
        while not bar.passed(parity):
            yield
        assert bar.phase == expect_phase + 1, f"{bar.name}: wait for phase {expect_phase} passed at phase {bar.phase}"

    def later(self, fn):
        self.pending_async.append(fn)

    # ---- agents ----
    def producer(self, r):
        c = self.c[r]
        for i in range(KB):
            j = order(i, r)
            c.a_ready[j].arrive_expect_tx(1)
            c.A_writers[j] += 1

            def land(j=j):
                assert c.A_readers[j] == 0
                c.A[j] = 0
                c.A_h0[j] = 0
                c.A_writers[j] -= 1
                c.a_ready[j].complete_tx(1)
            self.later(land)
            yield
        stage, phase, fills = 0, 0, 0
        for s in range(self.n):
            for i in range(KB):
                if fills >= BST:
                    yield from self.wait(c.b_empty[stage], phase ^ 1, fills // BST - 1)
                c.b_full[stage].arrive_expect_tx(1)

                def land(stage=stage, s=s, i=i):
                    c.B[stage] = (s, i)
                    c.b_full[stage].complete_tx(1)
                self.later(land)
                fills += 1
                stage += 1
                if stage == BST:
                    stage, phase = 0, phase ^ 1
                yield

    def retire_mma(self, r):
        """the tensor core retires batches in issue order; each carries the commits that follow it"""
        c = self.c[r]

        def fire():
            if not c.mma_queue:
                return
            batch = c.mma_queue.pop(0)
            for fn in batch:
                fn()
        return fire

    def mma(self, r):
        c = self.c[r]
        stage, phase, uses = 0, 0, 0
        for s in range(self.n):
            tb = s & 1
            for i in range(KB):
                j = order(i, r)
                halves = V3 and (not is_peer_slot(i)) and s > 0
                if halves:
                    b = j - r * 4
                    yield from self.wait(c.a_half[b], (s - 1) & 1, s - 1)
                else:
                    yield from self.wait(c.a_ready[j], s & 1, s)
                if is_peer_slot(i) and s + 1 < self.n:
                    c.a_ready[j].arrive_expect_tx(1)
                yield from self.wait(c.b_full[stage], phase, uses // BST)
                if halves:
                    # first two k-steps read the first half only
                    assert c.A_h0[j] == s, f"c{r} step {s}: first half of box {j} holds input of step {c.A_h0[j]}"
                    assert c.B[stage] == (s, i)
                    yield
                    yield from self.wait(c.a_ready[j], s & 1, s)
                # issue: checks at issue time
                assert c.A[j] == s, f"c{r} step {s}: box {j} holds input of step {c.A[j]}"
                assert c.A_writers[j] == 0, f"c{r} step {s}: box {j} is being written"
                assert c.B[stage] == (s, i), f"c{r}: weight stage {stage} holds {c.B[stage]}, want {(s, i)}"
                if i == 0:
                    assert c.tmem_reading[tb] == 0, f"c{r} step {s}: TMEM buffer {tb} still being drained"
                    assert s < 2 or c.drained.get(s - 2, 0) == 4, f"c{r} step {s}: epilogue {s-2} has not drained TMEM"
                    c.tmem[tb] = ["acc", s, False]
                c.A_readers[j] += 1
                fns = []

                def done(j=j, stage=stage, s=s, i=i, tb=tb):
                    c.A_readers[j] -= 1
                    c.b_empty[stage].arrive()
                    if i == KB - 1:
                        c.tmem[tb][2] = True
                        c.tmem_full[tb].arrive()
                fns.append(done)
                c.mma_queue.append(fns)
                self.later(self.retire_mma(r))
                uses += 1
                stage += 1
                if stage == BST:
                    stage, phase = 0, phase ^ 1
                yield
            # all MMAs of the step retired -> tell the peer its copies may overwrite this CTA's A buffer
            yield from self.wait(c.tmem_full[tb], (s >> 1) & 1, s >> 1)
            self.c[r ^ 1].peer_free.arrive()
            yield

    def epi(self, r, g):
        c, p = self.c[r], self.c[r ^ 1]
        groups = []  # this issuer's bulk store groups: each a dict(read_done=bool)
        for s in range(self.n):
            tb = s & 1
            last = s == self.n - 1
            yield from self.wait(c.tmem_full[tb], (s >> 1) & 1, s >> 1)
            yield from self.wait(c.peer_free, s & 1, s)  # (the group's issuer thread, before the group barrier)
            # bar.sync 3 (both groups)
            gen = self.bar3_gen[r]
            gen[g] += 1
            while gen[g ^ 1] < gen[g]:
                yield
            for box in (g, g + 2):
                j = r * 4 + box
                # issuer: cp.async.bulk.wait_group.read 1
                while sum(1 for q in groups[:-1] if not q["read_done"]) > 0:
                    yield
                yield
                assert c.tmem[tb] is not None and c.tmem[tb][1] == s and c.tmem[tb][2], f"c{r} step {s}: TMEM not ready"
                c.tmem_reading[tb] += 1
                yield
                c.tmem_reading[tb] -= 1
                c.drained[s] = c.drained.get(s, 0) + 1
                # write the own box
                assert c.A_readers[j] == 0, f"c{r} step {s}: overwriting box {j} with {c.A_readers[j]} readers in flight"
                assert c.A_writers[j] == 0
                if V3:
                    c.A_h0[j] = s + 1
                    yield
                    if not last:
                        c.a_half[box].arrive()
                    yield
                c.A[j] = s + 1
                c.A_h0[j] = s + 1
                yield
                if not last:
                    c.a_ready[j].arrive()
                    # DSMEM copy: reads my box, writes the peer's box, completes on the peer's barrier
                    c.A_readers[j] += 1
                    p.A_writers[j] += 1
                    assert p.A_readers[j] == 0, f"copy c{r}->c{r^1} step {s}: peer box {j} still read by its tensor core"

                    def land(j=j, s=s):
                        assert p.A_readers[j] == 0, f"copy landing: peer box {j} has readers"
                        p.A[j] = s + 1
                        p.A_h0[j] = s + 1
                        p.A_writers[j] -= 1
                        c.A_readers[j] -= 1
                        p.a_ready[j].complete_tx(1)
                    self.later(land)
                grp = {"read_done": False}
                groups.append(grp)
                c.A_readers[j] += 1

                def stored(j=j, s=s, grp=grp):
                    c.A_readers[j] -= 1
                    grp["read_done"] = True
                    c.stored.add((s, j))
                self.later(stored)
                yield

    def run(self, max_ticks=5_000_000):
        live = list(self.agents)
        for _ in range(max_ticks):
            if not live and not self.pending_async:
                break
            # choose between advancing an agent and firing an async completion
            if self.pending_async and (not live or self.rng.random() < self.p_async):
                k = self.rng.randrange(len(self.pending_async))
                # the tensor core retires in order: model by only allowing the retire callbacks (they pop the queue head)
                fn = self.pending_async.pop(k)
                fn()
                continue
            a = self.rng.choice(live)
            try:
                next(a)
            except StopIteration:
                live.remove(a)
        else:
            raise AssertionError("deadlock / livelock: agents did not finish")
        for r in (0, 1):
            want = {(s, r * 4 + b) for s in range(self.n) for b in range(4)}
            assert self.c[r].stored == want, f"c{r}: missing stores {sorted(want - self.c[r].stored)[:4]}"


def main():
    global V3
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    for V3 in (False, True):
        for seed in range(runs):
            for n in {n_steps, 1, 2, 3}:
                Sim(n, seed).run()
        print(f"ok ({'V3' if V3 else 'default'} protocol): {runs} random schedules x steps {{1,2,3,{n_steps}}}")


if __name__ == "__main__":
    main()
