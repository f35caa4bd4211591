"""TEST INFRASTRUCTURE: an executable model of the synchronisation protocol of motionbert_b200/csrc/mlp_fused.cuh.

One CTA pair of mlp_fused_kernel is restated as cooperating agents -- per CTA a TMA producer warp and 8 epilogue warps,
plus the leader's MMA issuer -- over modelled mbarriers (arrival counts, transaction counts, phase parity), an in-order
tensor pipe, asynchronous TMA loads and two-phase TMA stores (smem read, then global write) with bulk-group accounting.
Every agent follows the kernel's control flow statement by statement (same waits, same parities, same buffer choices:
`rc`, `hc`, `pend`, the deferred hidden-ready signal, the residual prefetch chain).  A randomised scheduler interleaves
the agents and draws every asynchronous latency; the model ASSERTS the data-flow properties the kernel relies on:

  * an MMA only consumes a stage that holds the operand tiles of exactly its (token block, tile, K block);
  * the fc2 mainloop only loads hidden columns whose stores of the SAME token block have completed (RAW through L2);
  * a hidden ring slot is only overwritten after every fc2 load of the previous token block has read it (WAR);
  * an accumulator is only overwritten after all 16 epilogue warps of the pair released it, and only read complete;
  * a staging buffer is only rewritten (by a thread, or by a residual load) after the TMA stores that read it did;
  * every output chunk is written exactly once, and the run terminates (no deadlock: some agent can always proceed).

tests/test_mlp_schedule_model.py runs it over tile geometries (base, Lite, hidden = 2048, one-tile shapes), block counts and
seeds.  It checks the PROTOCOL, not the arithmetic (tests/test_gpu_mlp_fused.py does that bit for bit on the GPU)."""
import random

STAGES = 4
EW = 8                      # epilogue warps per CTA


class Bar:
    """mbarrier: `count` pending arrivals + a transaction count; the phase flips when both reach zero."""

    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.phase ^= 1
            self.pending = self.count

    def arrive(self):
        assert self.pending > 0, "more arrivals than the barrier expects in one phase"
        self.pending -= 1
        self._check()

    def arrive_expect_tx(self, units):
        self.tx += units
        self.arrive()

    def complete_tx(self, units):
        self.tx -= units
        self._check()

    def passed(self, parity):            # mbarrier.try_wait.parity: the phase with this parity has completed
        return self.phase != parity


class Group:
    """one cp.async.bulk commit group of a thread: its stores, each read-done then write-done"""

    def __init__(self):
        self.stores = []                 # dicts {buf, read, done}

    @property
    def read_done(self):
        return all(s["read"] for s in self.stores)

    @property
    def done(self):
        return all(s["done"] for s in self.stores)


class Sim:
    def __init__(self, NT1, NT2, KB1, rounds, seed, ring=True, bug=None, store_read=(1, 10), store_write=(3, 60)):
        """bug: None (the kernel's protocol) or a deliberately broken variant the model must catch --
        'no_hready_wait' (producer reloads hidden columns without waiting for their flag), 'early_signal' (the flag is
        raised when the stores are issued, not when they completed), 'no_wait_read' (staging reused without waiting)."""
        self.NT1, self.NT2, self.KB1, self.KB2, self.rounds, self.ring = NT1, NT2, KB1, NT1 * 8, rounds, ring
        self.bug = bug
        self.store_read, self.store_write = store_read, store_write        # latency ranges of a TMA store's two phases (ticks)
        self.rng = random.Random(seed)
        self.t = 0
        self.events = []                 # (due, seq, fn)
        self.seq = 0
        # barriers
        self.full = [Bar(1) for _ in range(STAGES)]                                  # leader's
        self.empty = [[Bar(1) for _ in range(STAGES)] for _ in range(2)]
        self.tfull = [[Bar(1) for _ in range(2)] for _ in range(2)]
        self.tempty = [Bar(2 * EW) for _ in range(2)]                                # leader's
        self.rbar = [[[Bar(1) for _ in range(2)] for _ in range(EW)] for _ in range(2)]
        self.hready = [[Bar(EW) for _ in range(NT1)] for _ in range(2)]
        # storage
        self.stage = [[{"A": None, "B": None} for _ in range(STAGES)] for _ in range(2)]
        self.acc = [None, None]          # [k, e, kbs accumulated]
        self.acc_readers = [0, 0]
        self.hid = [dict() for _ in range(2)]      # per CTA: (n, half, ch, quad) -> {"k", "reads"}
        self.hid_inflight = [dict() for _ in range(2)]
        self.out = {}
        self.pipe = []                   # in-order tensor pipe: pending MMA ops
        self.pipe_busy = False

    # ------------------------------------------------------------ async machinery
    def later(self, lo, hi, fn):
        self.seq += 1
        self.events.append((self.t + self.rng.randint(lo, hi), self.seq, fn))

    def hidden_slot(self, k):
        return 0 if self.ring else k     # ring: ONE slot per pair, rewritten every token block

    def tma_load_operand(self, c, stage, which, tag, hidden_kb=None):
        k = tag[0]

        def land():
            if hidden_kb is not None:    # fc2 A operand: 128 rows x 32 hidden columns of THIS CTA
                n, half, ch = hidden_kb // 8, (hidden_kb % 8) // 4, hidden_kb % 4
                for quad in range(4):
                    key = (self.hidden_slot(k), n, half, ch, quad)
                    ent = self.hid[c].get(key)
                    assert ent is not None and ent["k"] == k, f"RAW: fc2 load of block {k} kb {hidden_kb} found {ent}"
                    ent["reads"] += 1
                    self.hid_inflight[c][key] -= 1
            assert self.stage[c][stage][which] is None, "stage overwritten before the MMA consumed it"
            self.stage[c][stage][which] = tag
            self.full[stage].complete_tx(1)
        if hidden_kb is not None:
            n, half, ch = hidden_kb // 8, (hidden_kb % 8) // 4, hidden_kb % 4
            for quad in range(4):
                key = (self.hidden_slot(k), n, half, ch, quad)
                self.hid_inflight[c][key] = self.hid_inflight[c].get(key, 0) + 1
        self.later(1, 30, land)

    def run_pipe(self):
        if self.pipe_busy or not self.pipe:
            return
        self.pipe_busy = True
        op = self.pipe.pop(0)

        def finish():
            stage, acc, k, e, kb, last = op
            for c in range(2):
                for which in "AB":
                    assert self.stage[c][stage][which] == (k, e, kb), \
                        f"MMA ({k},{e},{kb}) found {self.stage[c][stage][which]} in stage {stage} of CTA {c}"
                    self.stage[c][stage][which] = None
            if kb == 0:
                assert self.acc_readers[acc] == 0, "accumulator overwritten while epilogue warps still read it"
                self.acc[acc] = [k, e, 1]
            else:
                assert self.acc[acc][:2] == [k, e] and self.acc[acc][2] == kb
                self.acc[acc][2] += 1
            for c in range(2):           # tcgen05.commit ... multicast::cluster
                self.empty[c][stage].arrive()
                if last:
                    self.tfull[c][acc].arrive()
            self.pipe_busy = False
            self.run_pipe()
        self.later(1, 5, finish)

    # ------------------------------------------------------------ agents (generators; every yield is a preemption point)
    def producer(self, c):
        stage, phase = 0, 0
        for k in range(self.rounds):
            for e in range(self.NT1 + self.NT2):
                is_fc2 = e >= self.NT1
                n = e - self.NT1 if is_fc2 else e
                for kb in range(self.KB2 if is_fc2 else self.KB1):
                    if is_fc2 and n == 0 and kb % 8 == 0 and self.bug != "no_hready_wait":
                        yield ("wait", self.hready[c][kb // 8], k & 1)
                    yield ("wait", self.empty[c][stage], phase ^ 1)
                    if c == 0:
                        self.full[stage].arrive_expect_tx(4)             # 2 * STAGE_BYTES: A and B of both CTAs
                    self.tma_load_operand(c, stage, "A", (k, e, kb), hidden_kb=kb if is_fc2 else None)
                    self.tma_load_operand(c, stage, "B", (k, e, kb))
                    yield ("yield",)
                    stage += 1
                    if stage == STAGES:
                        stage, phase = 0, phase ^ 1

    def mma(self):
        stage, phase, acc, acc_phase = 0, 0, 0, 0
        for k in range(self.rounds):
            for e in range(self.NT1 + self.NT2):
                nkb = self.KB2 if e >= self.NT1 else self.KB1
                yield ("wait", self.tempty[acc], acc_phase ^ 1)
                for kb in range(nkb):
                    yield ("wait", self.full[stage], phase)
                    self.pipe.append((stage, acc, k, e, kb, kb == nkb - 1))
                    self.run_pipe()
                    yield ("yield",)
                    stage += 1
                    if stage == STAGES:
                        stage, phase = 0, phase ^ 1
                acc ^= 1
                if acc == 0:
                    acc_phase ^= 1

    def epilogue(self, c, w):
        quad, half = w % 4, w // 4
        groups = []                       # this warp's (elected lane's) bulk groups, oldest first
        bufs = {"buf0": None, "buf1": None, "bufS": None}        # staging content tags
        buf = ["buf0", "buf1"]
        my_rbar = self.rbar[c][w]

        def wait_read(n):                 # cp.async.bulk.wait_group.read n
            return lambda: all(g.read_done for g in (groups[:-n] if n else groups))

        def wait_done(n):                 # cp.async.bulk.wait_group n
            return lambda: all(g.done for g in (groups[:-n] if n else groups))

        def assert_free(name):
            for g in groups:
                for s in g.stores:
                    assert s["buf"] != name or s["read"], f"staging {name} rewritten while a TMA store still reads it"

        def store(group, name, on_done):
            s = {"buf": name, "read": False, "done": False}
            group.stores.append(s)

            def read():
                s["read"] = True

                def done():
                    s["done"] = True
                    on_done()
                self.later(self.store_write[0], self.store_write[1], done)
            self.later(self.store_read[0], self.store_read[1], read)

        def residual_load(b, tag):
            assert_free(buf[b])
            my_rbar[b].arrive_expect_tx(1)

            def land():
                bufs[buf[b]] = ("resid",) + tag
                my_rbar[b].complete_tx(1)
            self.later(1, 30, land)

        rc = hc = 0
        acc = acc_phase = 0
        pend = None
        for k in range(self.rounds):
            for e in range(self.NT1 + self.NT2):
                is_fc2 = e >= self.NT1
                n = e - self.NT1 if is_fc2 else e
                if not is_fc2:
                    yield ("wait", self.tfull[c][acc], acc_phase)
                    assert self.acc[acc] == [k, e, self.KB1], f"fc1 epilogue read an incomplete accumulator {self.acc[acc]}"
                    self.acc_readers[acc] += 1
                    for ch in range(4):
                        if ch == 1 and pend is not None:
                            if self.bug != "early_signal":
                                yield ("wait_fn", wait_done(1))
                            pend.arrive()
                        elif ch == 0:
                            yield ("wait_fn", wait_read(0))
                        elif self.bug != "no_wait_read":
                            yield ("wait_fn", wait_read(1))
                        if ch == 1:
                            pend = None
                        yield ("yield",)                                  # tcgen05.ld + math
                        if ch == 3:
                            self.acc_readers[acc] -= 1
                            self.tempty[acc].arrive()
                        ss = "bufS" if hc & 1 else buf[(rc & 1) ^ 1]
                        assert_free(ss)
                        bufs[ss] = ("hid", k, n, ch)
                        yield ("yield",)
                        key = (self.hidden_slot(k), n, half, ch, quad)
                        old = self.hid[c].get(key)
                        if old is not None and self.ring:
                            assert old["reads"] == self.NT2 and self.hid_inflight[c].get(key, 0) == 0, \
                                f"WAR: ring chunk {key} of block {old['k']} overwritten after {old['reads']} of {self.NT2} reads"
                        self.hid[c].pop(key, None)                       # content undefined until the store completes
                        g = Group()
                        store(g, ss, lambda key=key, k=k: self.hid[c].__setitem__(key, {"k": k, "reads": 0}))
                        groups.append(g)
                        hc += 1
                    pend = self.hready[c][n]
                    if n == self.NT1 - 1:
                        if self.bug != "early_signal":
                            yield ("wait_fn", wait_done(0))
                        pend.arrive()
                        pend = None
                else:
                    yield ("wait_fn", wait_read(0))
                    residual_load(rc & 1, (k, n, 0))
                    yield ("wait", self.tfull[c][acc], acc_phase)
                    assert self.acc[acc] == [k, e, self.KB2], f"fc2 epilogue read an incomplete accumulator {self.acc[acc]}"
                    self.acc_readers[acc] += 1
                    for ch in range(4):
                        b = rc & 1
                        yield ("wait", my_rbar[b], (rc >> 1) & 1)
                        assert bufs[buf[b]] == ("resid", k, n, ch), f"residual chunk mismatch: {bufs[buf[b]]} vs {(k, n, ch)}"
                        if ch < 3:
                            yield ("wait_fn", wait_read(0))
                            residual_load(b ^ 1, (k, n, ch + 1))
                        yield ("yield",)                                  # tcgen05.ld
                        if ch == 3:
                            self.acc_readers[acc] -= 1
                            self.tempty[acc].arrive()
                            yield ("wait_fn", wait_read(0))
                        yield ("yield",)                                  # math on the residual tile in buf[b]
                        assert_free(buf[b])
                        assert_free("bufS")
                        bufs[buf[b]] = ("out", k, n, ch)
                        bufs["bufS"] = ("outS", k, n, ch)
                        okey = (k, n, half, ch, quad, c)
                        g = Group()
                        store(g, buf[b], lambda okey=okey: self.out.__setitem__(okey, self.out.get(okey, 0) + 1))
                        store(g, "bufS", lambda: None)
                        groups.append(g)
                        rc += 1
                acc ^= 1
                if acc == 0:
                    acc_phase ^= 1
        yield ("wait_fn", wait_done(0))

    # ------------------------------------------------------------ scheduler
    def run(self, max_steps=2_000_000):
        agents = [self.producer(0), self.producer(1), self.mma()] + [self.epilogue(c, w) for c in range(2) for w in range(EW)]
        # scheduling bias per run: every agent gets its own speed (x 1/20 ... x 20), so that runs differ in WHICH role lags --
        # slow epilogue warps behind a fast MMA issuer is what exposes a missing hidden-ready wait, slow TMA what exposes ...
        speeds = [self.rng.choice([0.05, 0.3, 1.0, 1.0, 3.0, 20.0]) for _ in agents]
        cls = self.rng.choice([0.05, 1.0, 20.0])
        speeds[2] *= cls                                                   # the MMA issuer as a class of its own
        waiting = [("yield",)] * len(agents)
        alive = [True] * len(agents)
        for _ in range(max_steps):
            self.t += 1
            due = sorted(ev for ev in self.events if ev[0] <= self.t)
            self.events = [ev for ev in self.events if ev[0] > self.t]
            for ev in due:
                ev[2]()

            def runnable(i):
                w = waiting[i]
                return alive[i] and (w[0] == "yield" or (w[0] == "wait" and w[1].passed(w[2])) or (w[0] == "wait_fn" and w[1]()))
            ready = [i for i in range(len(agents)) if runnable(i)]
            if not ready:
                if not any(alive) and not self.events:
                    break
                assert self.events, f"DEADLOCK at t={self.t}: waiting {[(i, w[0]) for i, w in enumerate(waiting) if alive[i]]}"
                self.t = min(ev[0] for ev in self.events) - 1                # nothing to run: jump to the next completion
                continue
            i = self.rng.choices(ready, weights=[speeds[j] for j in ready])[0]
            try:
                waiting[i] = next(agents[i])
            except StopIteration:
                alive[i] = False
        else:
            raise AssertionError("model did not terminate")
        expected = {(k, n, half, ch, quad, c) for k in range(self.rounds) for n in range(self.NT2) for half in range(2)
                    for ch in range(4) for quad in range(4) for c in range(2)}
        assert set(self.out) == expected and all(v == 1 for v in self.out.values()), "every output chunk exactly once"
        return self.t
