"""Randomised model check of the cross-GPU protocol of the one-kernel data-parallel optimiser step (csrc/adamw_dp.cu,
adamw_dp_fused_kernel; no GPU needed).

G ranks, each a stream of iterations: COMPUTE (forward / backward: reads the rank's own weight shadows, overwrites its own gradient
buffer) then the optimiser kernel with NB blocks:
  1  block 0: fence, then "gradient complete" signal (epoch e) to every rank;  every block waits for every rank's signal
  2  every block reads its part of EVERY rank's gradient
  3  block arrives on a counter; the last one sends this rank's verdict (epoch-stamped) to every rank
  4  every block waits for every rank's verdict
  5  every block pushes its part of the new weights into EVERY rank's shadow; fence (its pushes have landed); arrives on counter 2
  6  the last block: epoch += 1, "weights written" signal to every rank, waits for everybody's; kernel complete
Remote stores (signals and weight pushes) are delivered after random delays, in any order, except that a fence waits for the
issuing agent's own earlier stores. Agents of all ranks are interleaved at random. Asserted at every step:
  * a gradient is read only while it is the gradient of THIS iteration (not overwritten early by the owner's next COMPUTE, not stale),
  * a weight shadow is overwritten only while its owner is not computing with it, and COMPUTE of iteration i + 1 sees the weights
    of iteration i from EVERY rank (all pushes landed),
  * a signal wait never passes on an older or aliased epoch, the verdict every rank acts on is the OR of all ranks' verdicts of
    this iteration,
  * no deadlock.
Run: python tools/sim_dp_protocol.py [runs] [ranks] [iterations]
"""
import random
import sys


class Rank:
    def __init__(self, g, G, NB):
        self.g = g
        self.sig = {"grads": [0] * G, "verdict": [0] * G, "applied": [0] * G}   # written by the peers, polled locally
        self.epoch = 0                      # device-resident: iterations completed
        self.counter = [0, 0]
        self.bad_bits = 0
        self.grad_version = 0               # iteration whose gradient the buffer holds (0: none)
        self.grad_readers = 0
        self.weights_from = [0] * G         # shadow part owned by rank q holds the weights of iteration weights_from[q]
        self.computing = False
        self.iteration_done = 0


class Sim:
    def __init__(self, G, NB, iters, seed):
        self.G, self.NB, self.iters = G, NB, iters
        self.rng = random.Random(seed)
        self.p_deliver = (0.02, 0.1, 0.4, 0.9)[seed % 4]
        self.ranks = [Rank(g, G, NB) for g in range(G)]
        self.inflight = []                  # (owner agent id, fn)
        self.bad_truth = {}                 # (iteration) -> per-rank bad flags drawn at random
        self.acted = {}                     # (iteration, rank) -> verdict the rank acted on
        self.agents = []
        for g in range(G):
            self.agents.append(self.stream(g))

    # remote store: delivered later, unordered; `tag` identifies the issuing agent for fences
    def remote(self, tag, fn):
        self.inflight.append((tag, fn))

    def fence(self, tag):
        while any(t == tag for t, _ in self.inflight):
            yield

    def deliver_some(self):
        i = 0
        while i < len(self.inflight):
            if self.rng.random() < self.p_deliver:
                _, fn = self.inflight.pop(i)
                fn()
            else:
                i += 1

    def stream(self, g):
        me = self.ranks[g]
        for it in range(1, self.iters + 1):
            # ---- COMPUTE: forward / backward of iteration `it` ----
            me.computing = True
            assert all(v == it - 1 for v in me.weights_from), f"rank {g} computes iteration {it} with weights {me.weights_from}"
            yield
            assert me.grad_readers == 0, f"rank {g} overwrites its gradient under {me.grad_readers} readers"
            me.grad_version = it
            yield
            assert all(v == it - 1 for v in me.weights_from), f"rank {g}: weights changed during the compute of iteration {it}"
            me.computing = False
            bad = self.bad_truth.setdefault(it, [self.rng.random() < 0.15 for _ in range(self.G)])
            # ---- the optimiser kernel: NB blocks as sub-agents, interleaved with everything else ----
            blocks = [self.block(g, b, it, bad[g]) for b in range(self.NB)]
            live = list(blocks)
            while live:
                b = self.rng.choice(live)
                try:
                    next(b)
                except StopIteration:
                    live.remove(b)
                yield
            assert me.iteration_done == it

    def wait_row(self, me, row, q, e):
        while me.sig[row][q] < e:
            yield
        assert me.sig[row][q] == e or row != "verdict", f"rank {me.g}: {row} signal of rank {q} is {me.sig[row][q]}, waited for {e}"

    def block(self, g, b, it, my_bad):
        me, G = self.ranks[g], self.G
        tag = (g, b, it)
        e = me.epoch + 1
        assert e == it, f"rank {g} block {b}: epoch {me.epoch} at iteration {it}"
        if b == 0:                                                   # 1: gradient complete
            yield from self.fence(tag)
            for q in range(G):
                self.remote(tag, lambda q=q: self.ranks[q].sig["grads"].__setitem__(g, max(self.ranks[q].sig["grads"][g], e)))
        for q in range(G):
            yield from self.wait_row(me, "grads", q, e)
        for q in range(G):                                           # 2: read every rank's gradient (its own part of it)
            peer = self.ranks[q]
            assert peer.grad_version == it, f"rank {g} reads rank {q}'s gradient of iteration {peer.grad_version} in iteration {it}"
            peer.grad_readers += 1
            yield
            assert peer.grad_version == it, f"rank {q}'s gradient was overwritten under rank {g}'s read"
            peer.grad_readers -= 1
        if my_bad and b == 0:                                        # 3: verdict
            me.bad_bits |= 1
        me.counter[0] += 1
        if me.counter[0] == self.NB:
            me.counter[0] = 0
            v = e * 2 + (me.bad_bits & 1)
            me.bad_bits = 0
            for q in range(G):
                self.remote(tag, lambda q=q, v=v: self.ranks[q].sig["verdict"].__setitem__(g, v))
        found = 0
        for q in range(G):                                           # 4: everybody's verdict
            while (me.sig["verdict"][q] >> 1) < e:
                yield
            assert (me.sig["verdict"][q] >> 1) == e, f"rank {g}: verdict of rank {q} from epoch {me.sig['verdict'][q] >> 1} in epoch {e}"
            found |= me.sig["verdict"][q] & 1
        truth = int(any(self.bad_truth[it]))
        assert found == truth, f"rank {g} block {b}: acts on verdict {found}, truth {truth}"
        self.acted[(it, g, b)] = found
        for q in range(G):                                           # 5: push this block's part of the new weights
            def land(q=q):
                peer = self.ranks[q]
                assert not peer.computing, f"rank {g} writes rank {q}'s weights while it computes"
                peer._parts = getattr(peer, "_parts", {})
                k = (it, g)
                peer._parts[k] = peer._parts.get(k, 0) + 1
                if peer._parts[k] == self.NB:
                    peer.weights_from[g] = it                        # (a skipped step 'pushes' the unchanged weights: same protocol)
            self.remote(tag, land)
        yield from self.fence(tag)
        me.counter[1] += 1
        if me.counter[1] == self.NB:                                 # 6: last block
            me.counter[1] = 0
            me.epoch = e
            yield from self.fence(tag)
            for q in range(G):
                self.remote(tag, lambda q=q: self.ranks[q].sig["applied"].__setitem__(g, max(self.ranks[q].sig["applied"][g], e)))
            for q in range(G):
                yield from self.wait_row(me, "applied", q, e)
            me.iteration_done = it

    def run(self):
        live = list(self.agents)
        idle = 0
        while live:
            a = self.rng.choice(live)
            before = (len(self.inflight), tuple(r.epoch for r in self.ranks), tuple(tuple(r.sig[k]) for r in self.ranks for k in r.sig))
            try:
                next(a)
            except StopIteration:
                live.remove(a)
            self.deliver_some()
            after = (len(self.inflight), tuple(r.epoch for r in self.ranks), tuple(tuple(r.sig[k]) for r in self.ranks for k in r.sig))
            idle = idle + 1 if before == after else 0
            assert idle < 200000, "no progress: deadlock"
        while self.inflight:
            self.deliver_some()
        assert all(r.iteration_done == self.iters for r in self.ranks)


if __name__ == "__main__":
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    G = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    for seed in range(runs):
        Sim(G, 3, iters, seed).run()
    print(f"{runs} random interleavings of {G} ranks x 3 blocks x {iters} iterations: protocol holds")
