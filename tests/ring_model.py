"""Executable model of the flow control of the production kernel (word2bits_b200/csrc/w2b_ring.cuh).

TEST INFRASTRUCTURE.  The TMA ring kernel runs one sampler warp, one loader warp and `ncw` consumer
warps per CTA that hand shared-memory row slots to each other through mbarriers ("rows landed"),
per-slot release counters ("slot free again") and monotonic counters (descriptors, u-ring space).  A
mistake in that protocol shows up on the GPU as a hang (one did: DESIGN.md section 4.1) or as a row
overwritten while the TMA still reads it.  This module restates the protocol — every wait, every
counter update, every asynchronous copy, in the order the kernel performs them — as cooperating
generators driven by a randomised scheduler, so that the geometry chosen by the host planner
(`w2b_ring_plan_query`) can be checked on a CPU for

  * liveness: every schedule ends with all positions trained (no wait can block forever);
  * safety: a slot is never loaded into, overwritten or released while a consumer still has to read
    it or a bulk reduce has not confirmed reading it; an mbarrier is never re-armed while a consumer
    may still wait for its previous phase; staging rows and descriptors are not reused early.

Asynchronous operations are adversarial: a bulk load lands at an arbitrary later scheduling point,
a bulk reduce finishes reading its source only when a `cp.async.bulk.wait_group.read` forces it.

The model follows the kernel's names (nu, nv, G, R, kND, prog, urel, s_rc, prev[], pend_u ...).  It
describes the default (prefetching) mode; `ring_serial` is the same protocol with extra waits.
"""
import random


class Deadlock(AssertionError):
    pass


class Barrier:
    """mbarrier used with one arrival per phase + transaction bytes (rows here)."""

    def __init__(self):
        self.completed = 0   # phases completed
        self.pending = None  # rows still to land in the armed phase (None = not armed)

    def arm(self, rows):
        assert self.pending is None, "mbarrier re-armed before its phase completed"
        if rows == 0:
            self.completed += 1
        else:
            self.pending = rows

    def land(self):
        self.pending -= 1
        if self.pending == 0:
            self.pending = None
            self.completed += 1


class BulkQueue:
    """Per-thread bulk async-group queue (cp.async.bulk.commit_group / wait_group.read N)."""

    def __init__(self):
        self.open = []    # sources of reduces issued since the last commit
        self.groups = []  # committed groups whose reads are not confirmed yet

    def issue(self, src):
        src["reading"] = src.get("reading", 0) + 1
        self.open.append(src)

    def commit(self):
        self.groups.append(self.open)
        self.open = []

    def wait_read(self, n):
        while len(self.groups) > n:
            for src in self.groups.pop(0):
                src["reading"] -= 1


class RingModel:
    def __init__(self, plan, window, negative, positions, seed=0, kND=None, early_release=False):
        self.early_release = early_release  # variant kernels, TrainParams::serial == 2
        self.nu, self.nv = plan["u_rows"], plan["v_rows"]
        self.G, self.R, self.ncw = plan["group"], plan["rows_in_flight"], plan["consumer_warps"]
        self.upw = plan.get("units_per_warp", 1) or 1  # row units per consumer warp (LPR < 32 variants)
        self.kND = kND or plan["desc_depth"]
        self.kMaxGrp = plan["max_groups"]
        self.ngmax = (negative + 1 + self.G - 1) // self.G
        assert self.ngmax <= self.kMaxGrp, "more groups per position than landing barriers"
        assert self.G <= 32, "a group is loaded by the lanes of one warp"
        self.window, self.negative = window, negative
        self.pos = positions  # list of (cw, nt)
        for cw, nt in positions:
            assert 1 <= cw <= 2 * window and 1 <= nt <= negative + 1
        self.rng = random.Random(seed)
        # shared state (RingCtl + the rings)
        self.desc_ready = 0
        self.prog = 0
        self.urel = 0
        self.s_rc = [0] * self.nv
        self.ubar = [Barrier() for _ in range(self.kND)]
        self.vbar = [[Barrier() for _ in range(self.kMaxGrp)] for _ in range(self.kND)]
        self.desc = [None] * self.kND           # descriptor slot -> dict(q, cw, nt, us0, vs0, exit)
        self.uslot = [dict(state="free", row=None) for _ in range(self.nu)]
        self.vslot = [dict(state="free", row=None) for _ in range(self.nv)]
        self.errbuf = [dict(), dict()]          # staging rows of the u scatter
        self.inflight = []                      # bulk loads issued, not landed: (barrier, slot dict)
        self.bar_arrivals = {"A": 0, "B": 0}
        self.ctx_consumed = set()               # positions whose context rows every warp has read
        self.trained = [0] * len(positions)     # target rows processed per position
        self.u_scattered = set()
        self.finished = 0

    # ------------------------------------------------------------------ sampler warp
    def sampler(self):
        q = 0
        for q, (cw, nt) in enumerate(self.pos):
            yield lambda q=q: q - self.prog < self.kND
            slot = q % self.kND
            old = self.desc[slot]
            assert old is None or old["released"], "descriptor slot reused while still referenced"
            self.desc[slot] = dict(q=q, cw=cw, nt=nt, us0=None, vs0=None, exit=False, released=False)
            self.desc_ready = q + 1
        n = len(self.pos)
        yield lambda: n - self.prog < self.kND
        old = self.desc[n % self.kND]
        assert old is None or old["released"], "descriptor slot reused while still referenced"
        self.desc[n % self.kND] = dict(q=n, cw=0, nt=0, exit=True, released=True)
        self.desc_ready = n + 1
        self.finished += 1

    # ------------------------------------------------------------------- loader warp
    def _load(self, bar, slot, row):
        assert slot["state"] == "free", "bulk load into a slot that is still in use (%s)" % slot["state"]
        assert slot.get("reading", 0) == 0, "bulk load into a slot a bulk reduce is still reading"
        slot["state"], slot["row"] = "loading", row
        self.inflight.append((bar, slot))

    def loader(self):
        u_alloc = v_alloc = 0
        q = 0
        while True:
            yield lambda q=q: self.desc_ready > q
            slot = q % self.kND
            d = self.desc[slot]
            assert d["q"] == q
            if d["exit"]:
                self.ubar[slot].arm(0)
                break
            cw, nt = d["cw"], d["nt"]
            yield lambda ua=u_alloc, cw=cw: ua + cw - self.urel <= self.nu
            d["us0"], d["vs0"] = u_alloc % self.nu, v_alloc % self.nv
            self.ubar[slot].arm(cw)
            for k in range(cw):
                s = self.uslot[(u_alloc + k) % self.nu]
                if s["row"] is not None:
                    assert s["row"][0] in self.ctx_consumed, "context slot reloaded before it was read"
                    s["state"] = "free"
                self._load(self.ubar[slot], s, (q, k))
            u_alloc += cw
            for gi in range(self.ngmax):
                g0 = gi * self.G
                ng = max(0, min(self.G, nt - g0))
                bar = self.vbar[slot][gi]
                bar.arm(ng)
                # every lane waits for its own slot; lanes are independent threads
                lanes = list(range(ng))
                while lanes:
                    yield lambda lanes=lanes, va=v_alloc: any(
                        self.s_rc[(va + l) % self.nv] >= (va + l) // self.nv for l in lanes)
                    for l in list(lanes):
                        vi = v_alloc + l
                        if self.s_rc[vi % self.nv] >= vi // self.nv:
                            self._load(bar, self.vslot[vi % self.nv], (q, g0 + l))
                            lanes.remove(l)
                v_alloc += ng
            q += 1
        self.finished += 1

    # ---------------------------------------------------------------- consumer warps
    def _cta_barrier(self, name):
        self.bar_arrivals[name] += 1
        n = self.bar_arrivals[name]
        gen = (n - 1) // self.ncw
        return lambda: self.bar_arrivals[name] >= (gen + 1) * self.ncw

    def consumer(self, warp):
        ncw, R, G, nv, kND, upw = self.ncw, self.R, self.G, self.nv, self.kND, self.upw
        nunits = ncw * upw
        issuer = warp == ncw - 1
        # one bulk-group queue and one prev[] per unit leader; lane 0 (the issuer's thread) leads unit 0
        units = [dict(bq=BulkQueue(), prev=[-1] * R) for _ in range(upw)]
        pend = None               # issuer: (descriptor, q) whose u scatter is staged
        q = 0
        while True:
            slot = q % kND
            phase = q // kND
            yield lambda slot=slot, phase=phase: self.ubar[slot].completed >= phase + 1
            assert self.ubar[slot].completed == phase + 1, "u barrier ran a phase ahead of a waiting consumer"
            d = self.desc[slot]
            assert d["q"] == q, "descriptor overwritten before the consumers read it"
            fin = d["exit"]
            if not fin:
                for k in range(d["cw"]):   # context phase: reads the landed u rows
                    s = self.uslot[(d["us0"] + k) % self.nu]
                    assert s["state"] == "full" and s["row"] == (q, k), "context row not in its slot"
            yield self._cta_barrier("A")
            if not fin:
                self.ctx_consumed.add(q)
            if issuer:
                if not fin:
                    self.urel += d["cw"]
                if pend is not None:
                    pd, pq = pend
                    assert self.desc[pq % kND] is pd, "descriptor of the pending u scatter was recycled"
                    eb = self.errbuf[pq & 1]
                    assert eb.get("holds") == pq, "staging row does not hold this position's error"
                    for _ in range(pd["cw"]):
                        units[0]["bq"].issue(eb)
                    units[0]["bq"].commit()
                    self.u_scattered.add(pq)
                    pd["released"] = True
                    pend = None
                    self.prog = pq + 1
            if fin:
                break
            nt, vs0 = d["nt"], d["vs0"]
            i0w = warp * upw
            while i0w < nt:
                if self.early_release:  # leaders confirm the previous pass's reduces and free its slots first
                    for u in units:
                        if any(x >= 0 for x in u["prev"]):
                            u["bq"].wait_read(0)
                            for t in range(R):
                                if u["prev"][t] >= 0:
                                    self._release(u["prev"][t])
                                    u["prev"][t] = -1
                batch = []
                for sub in range(upw):
                    i0 = i0w + sub
                    ifall = i0 if (upw == 1 or i0 < nt) else i0w
                    rows = [i0 + t * nunits for t in range(R)]
                    have = [i < nt for i in rows]
                    batch.append((rows, have, ifall))
                # every lane waits for the barrier of the row it is going to read (LPR < 32: also the
                # lanes without a row of their own, on the landed row they re-read)
                for rows, have, ifall in batch:
                    for t in range(R):
                        if have[t] or upw > 1:
                            gi = (rows[t] if have[t] else ifall) // G
                            yield lambda slot=slot, gi=gi, phase=phase: self.vbar[slot][gi].completed >= phase + 1
                            assert self.vbar[slot][gi].completed == phase + 1, "v barrier ran a phase ahead"
                for u, (rows, have, ifall) in zip(units, batch):
                    sl = [(vs0 + (rows[t] if have[t] else ifall)) % nv for t in range(R)]
                    for t in range(R):
                        s = self.vslot[sl[t]]
                        if have[t]:
                            assert s["state"] == "full" and s["row"] == (q, rows[t]), \
                                "target row (%d,%d) not in slot %d: %s" % (q, rows[t], sl[t], s)
                        else:  # re-read of a landed row of this position (gets g = 0)
                            assert s["row"] == (q, ifall) and s["state"] in ("full", "update"), \
                                "fallback row (%d,%d) not readable in slot %d: %s" % (q, ifall, sl[t], s)
                    for t in range(R):
                        if have[t]:
                            self.vslot[sl[t]]["state"] = "update"   # overwritten in place with g * context_avg
                            self.trained[q] += 1
                    for t in range(R):
                        if have[t]:
                            u["bq"].issue(self.vslot[sl[t]])
                    u["bq"].commit()
                    u["bq"].wait_read(1)
                    for t in range(R):
                        if u["prev"][t] >= 0:
                            self._release(u["prev"][t])
                        u["prev"][t] = sl[t] if have[t] else -1
                i0w += R * nunits
                yield lambda: True  # a scheduling point between batches
            for u in units:
                u["bq"].wait_read(0)
                for t in range(R):
                    if u["prev"][t] >= 0:
                        self._release(u["prev"][t])
                    u["prev"][t] = -1
            yield self._cta_barrier("B")
            if warp == 0:  # (thread-per-column in the kernel) partial sums -> staging row q & 1
                eb = self.errbuf[q & 1]
                assert eb.get("reading", 0) == 0, "staging row rewritten while its scatter is still reading it"
                eb["holds"] = q
            if issuer:
                pend = (d, q)
            q += 1
        for u in units:
            u["bq"].wait_read(0)
        self.finished += 1

    def _release(self, sl):
        s = self.vslot[sl]
        assert s["state"] == "update" and s.get("reading", 0) == 0, "slot released before its reduce was read"
        s["state"] = "free"
        self.s_rc[sl] += 1

    # ---------------------------------------------------------------------- scheduler
    def run(self, max_steps=10_000_000):
        actors = [self.sampler(), self.loader()] + [self.consumer(w) for w in range(self.ncw)]
        waiting = {}
        for a in actors:
            waiting[a] = next(a)
        steps = 0
        while waiting:
            steps += 1
            assert steps < max_steps, "model did not terminate"
            # asynchronous loads land at arbitrary times
            if self.inflight and self.rng.random() < 0.3:
                self._land(self.rng.randrange(len(self.inflight)))
            ready = [a for a, cond in waiting.items() if cond()]
            if not ready:
                if self.inflight:
                    self._land(self.rng.randrange(len(self.inflight)))
                    continue
                raise Deadlock("no warp can make progress: desc_ready=%d prog=%d urel=%d trained=%s" % (
                    self.desc_ready, self.prog, self.urel, self.trained[:self.prog + 2]))
            a = self.rng.choice(ready)
            try:
                waiting[a] = next(a)
            except StopIteration:
                del waiting[a]
        assert self.finished == self.ncw + 2
        assert self.trained == [nt for _, nt in self.pos], "not every target row was processed exactly once"
        assert self.u_scattered == set(range(len(self.pos))), "a context scatter was lost"
        assert not self.inflight
        return steps

    def _land(self, i):
        bar, slot = self.inflight.pop(i)
        slot["state"] = "full"
        bar.land()


def random_positions(rng, n, window, negative, style):
    """cw in [1, 2*window], nt in [1, 1+negative]; `style` picks the mix."""
    out = []
    for _ in range(n):
        if style == "typical":      # interior of a sentence, almost no skipped negatives
            cw = rng.randint(max(1, window), 2 * window)
            nt = negative + 1 if rng.random() < 0.9 else rng.randint(max(1, negative - 1), negative + 1)
        elif style == "extreme":    # everything at its maximum
            cw, nt = 2 * window, negative + 1
        elif style == "tiny":       # sentence edges, heavy skipping
            cw, nt = rng.randint(1, min(2, 2 * window)), rng.randint(1, min(3, negative + 1))
        else:                       # anything goes
            cw, nt = rng.randint(1, 2 * window), rng.randint(1, negative + 1)
        out.append((cw, nt))
    return out
