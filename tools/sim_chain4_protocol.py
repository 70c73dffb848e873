"""Randomised model check of the synchronisation protocol of acezero_b200/csrc/head_chain4.cu (cluster of 4 CTAs,
tcgen05 cta_group::2; no GPU needed). Same method as tools/sim_chain_protocol.py.

CTA rank = 2 c + r: pair P_c = {2c, 2c+1} shares every weight k-block and runs ONE UMMA stream issued by its leader (r = 0)
that reads the A buffers of BOTH CTAs; the other CTA's warp 1 relays "my k-block j is in place" to the leader's
partner_ready[j]; boxes are exchanged with rank ^ 2 (same row tile, other channel half). Asserted at every step:
  * no mbarrier wait passes on a stale / aliased phase,
  * a UMMA of step s is issued only while k-block j of BOTH CTAs of the pair holds the input of step s and both halves of the
    weight stage hold (s, i),
  * nobody writes a shared-memory box under an in-flight reader (tensor core, DSMEM copy, TMA store),
  * TMEM buffers are drained before they are overwritten and complete before they are read,
  * no deadlock; every tile is stored.
Run: python tools/sim_chain4_protocol.py [runs] [n_steps]
"""
import random
import sys

KB, BST = 8, 6


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
        self.rank, self.c, self.r = rank, rank >> 1, rank & 1
        n = f"c{rank}"
        self.a_ready = [MBar(f"{n}.a_ready{j}") for j in range(KB)]
        self.partner_ready = [MBar(f"{n}.partner_ready{j}") for j in range(KB)]
        self.b_full = [MBar(f"{n}.b_full{i}", 2) for i in range(BST)]
        self.b_empty = [MBar(f"{n}.b_empty{i}") for i in range(BST)]
        self.tmem_full = [MBar(f"{n}.tmem_full{i}") for i in range(2)]
        self.peer_free = MBar(f"{n}.peer_free")
        self.A = [None] * KB
        self.A_readers = [0] * KB
        self.A_writers = [0] * KB
        self.B = [None] * BST          # this CTA's half of the weight stage: (step, i)
        self.tmem = [None, None]
        self.tmem_reading = [0, 0]
        self.drained = {}
        self.stored = set()


def order(i, c):
    b = ((i >> 2) << 1) | (i & 1)
    return ((c ^ 1) if (i & 2) else c) * 4 + b


class Sim:
    def __init__(self, n_steps, seed):
        self.n = n_steps
        self.rng = random.Random(seed)
        self.p_async = (0.01, 0.05, 0.35, 0.7)[seed % 4]
        self.cta = [CTA(k) for k in range(4)]
        self.pending_async = []
        self.mma_queue = {0: [], 1: []}   # per pair: issued, not yet retired UMMA batches (retire in order)
        self.bar3 = [[0, 0] for _ in range(4)]
        self.agents = []
        for k in range(4):
            self.agents += [self.producer(k), self.warp1(k), self.epi(k, 0), self.epi(k, 1)]

    def wait(self, bar, parity, expect_phase):
        while not bar.passed(parity):
            yield
        assert bar.phase == expect_phase + 1, f"{bar.name}: wait for phase {expect_phase} passed at phase {bar.phase}"

    def later(self, fn):
        self.pending_async.append(fn)

    # ------------------------------------------------------------------ producer (every CTA)
    def producer(self, k):
        me = self.cta[k]
        lead = self.cta[k & ~1]
        for i in range(KB):
            j = order(i, me.c)
            me.a_ready[j].arrive_expect_tx(1)
            me.A_writers[j] += 1

            def land(j=j):
                assert me.A_readers[j] == 0
                me.A[j] = 0
                me.A_writers[j] -= 1
                me.a_ready[j].complete_tx(1)
            self.later(land)
            yield
        stage, phase, fills = 0, 0, 0
        for s in range(self.n):
            for i in range(KB):
                if fills >= BST:
                    yield from self.wait(me.b_empty[stage], phase ^ 1, fills // BST - 1)
                if me.r == 0:
                    lead.b_full[stage].arrive_expect_tx(2)     # both halves
                else:
                    lead.b_full[stage].arrive()

                def land(stage=stage, s=s, i=i):
                    me.B[stage] = (s, i)
                    lead.b_full[stage].complete_tx(1)
                self.later(land)
                fills += 1
                stage += 1
                if stage == BST:
                    stage, phase = 0, phase ^ 1
                yield

    def retire(self, pair):
        def fire():
            q = self.mma_queue[pair]
            if q:
                for fn in q.pop(0):
                    fn()
        return fire

    # ------------------------------------------------------------------ warp 1: UMMA issuer (leader) / relay (partner)
    def warp1(self, k):
        me = self.cta[k]
        xp = self.cta[k ^ 2]
        if me.r == 0:
            partner = self.cta[k | 1]
            both = (me, partner)
            stage, phase, uses = 0, 0, 0
            for s in range(self.n):
                tb = s & 1
                for i in range(KB):
                    j = order(i, me.c)
                    yield from self.wait(me.a_ready[j], s & 1, s)
                    if (i & 2) and s + 1 < self.n:
                        me.a_ready[j].arrive_expect_tx(1)
                    yield from self.wait(me.partner_ready[j], s & 1, s)
                    yield from self.wait(me.b_full[stage], phase, uses // BST)
                    for x in both:
                        assert x.A[j] == s, f"pair {me.c} step {s}: c{x.rank} box {j} holds input of step {x.A[j]}"
                        assert x.A_writers[j] == 0, f"pair {me.c} step {s}: c{x.rank} box {j} is being written"
                        assert x.B[stage] == (s, i), f"pair {me.c}: c{x.rank} weight stage {stage} holds {x.B[stage]}, want {(s, i)}"
                        if i == 0:
                            assert x.tmem_reading[tb] == 0
                            assert s < 2 or x.drained.get(s - 2, 0) == 4, f"c{x.rank} step {s}: epilogue {s-2} has not drained TMEM"
                            x.tmem[tb] = ["acc", s, False]
                        x.A_readers[j] += 1

                    def done(j=j, stage=stage, i=i, tb=tb):
                        for x in both:
                            x.A_readers[j] -= 1
                            x.b_empty[stage].arrive()           # multicast commit
                            if i == KB - 1:
                                x.tmem[tb][2] = True
                                x.tmem_full[tb].arrive()
                    self.mma_queue[me.c].append([done])
                    self.later(self.retire(me.c))
                    uses += 1
                    stage += 1
                    if stage == BST:
                        stage, phase = 0, phase ^ 1
                    yield
                yield from self.wait(me.tmem_full[tb], (s >> 1) & 1, s >> 1)
                xp.peer_free.arrive()
                yield
        else:
            lead = self.cta[k & ~1]
            for s in range(self.n):
                tb = s & 1
                for i in range(KB):
                    j = order(i, me.c)
                    yield from self.wait(me.a_ready[j], s & 1, s)
                    if (i & 2) and s + 1 < self.n:
                        me.a_ready[j].arrive_expect_tx(1)
                    lead.partner_ready[j].arrive()
                    yield
                yield from self.wait(me.tmem_full[tb], (s >> 1) & 1, s >> 1)
                xp.peer_free.arrive()
                yield

    # ------------------------------------------------------------------ epilogue group g of CTA k
    def epi(self, k, g):
        me, xp = self.cta[k], self.cta[k ^ 2]
        groups = []
        for s in range(self.n):
            tb = s & 1
            last = s == self.n - 1
            yield from self.wait(me.tmem_full[tb], (s >> 1) & 1, s >> 1)
            yield from self.wait(me.peer_free, s & 1, s)
            gen = self.bar3[k]
            gen[g] += 1
            while gen[g ^ 1] < gen[g]:
                yield
            for box in (g, g + 2):
                j = me.c * 4 + box
                while sum(1 for q in groups[:-1] if not q["read_done"]) > 0:
                    yield
                yield
                assert me.tmem[tb] is not None and me.tmem[tb][1] == s and me.tmem[tb][2], f"c{k} step {s}: TMEM not ready"
                me.tmem_reading[tb] += 1
                yield
                me.tmem_reading[tb] -= 1
                me.drained[s] = me.drained.get(s, 0) + 1
                assert me.A_readers[j] == 0, f"c{k} step {s}: overwriting box {j} with {me.A_readers[j]} readers in flight"
                assert me.A_writers[j] == 0
                me.A[j] = s + 1
                yield
                if not last:
                    me.a_ready[j].arrive()
                    me.A_readers[j] += 1
                    xp.A_writers[j] += 1
                    assert xp.A_readers[j] == 0, f"copy c{k}->c{k^2} step {s}: box {j} of the target still read by its tensor core"

                    def land(j=j, s=s):
                        assert xp.A_readers[j] == 0
                        xp.A[j] = s + 1
                        xp.A_writers[j] -= 1
                        me.A_readers[j] -= 1
                        xp.a_ready[j].complete_tx(1)
                    self.later(land)
                grp = {"read_done": False}
                groups.append(grp)
                me.A_readers[j] += 1

                def stored(j=j, s=s, grp=grp):
                    me.A_readers[j] -= 1
                    grp["read_done"] = True
                    me.stored.add((s, j))
                self.later(stored)
                yield

    def run(self, max_ticks=8_000_000):
        live = list(self.agents)
        for _ in range(max_ticks):
            if not live and not self.pending_async:
                break
            if self.pending_async and (not live or self.rng.random() < self.p_async):
                self.pending_async.pop(self.rng.randrange(len(self.pending_async)))()
                continue
            a = self.rng.choice(live)
            try:
                next(a)
            except StopIteration:
                live.remove(a)
        else:
            raise AssertionError("deadlock / livelock: agents did not finish")
        for x in self.cta:
            want = {(s, x.c * 4 + b) for s in range(self.n) for b in range(4)}
            assert x.stored == want, f"c{x.rank}: missing stores {sorted(want - x.stored)[:4]}"


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    for seed in range(runs):
        for n in {n_steps, 1, 2, 3}:
            Sim(n, seed).run()
    print(f"ok (cluster-of-4 / cta_group::2 protocol): {runs} random schedules x steps {{1,2,3,{n_steps}}}")


if __name__ == "__main__":
    main()
