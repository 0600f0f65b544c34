"""Discrete-event model of the barrier protocol of csrc/attention_tc2.cu (two query tiles per CTA), runnable on a CPU.

The kernel was written without GPU time; this model executes ONE CTA's warp roles (TMA producer, MMA issuer, 8 softmax
warps) as coroutines over mbarriers with the hardware's phase-PARITY semantics, asynchronous TMA completions and
in-order asynchronous tensor-core completions (tcgen05.commit arrives when everything issued before it has finished),
under a random scheduler, and checks every shared resource for hazards:

  * info slot / Q tiles republished only after every reader is done with them,
  * K / V ring slots overwritten only after the MMAs that read them have completed, and read only when they hold the
    expected key block,
  * S_t overwritten only after the tile's four softmax warps have loaded the previous S_t,
  * P_t overwritten only after the P V product that reads it has completed; P V issued only on a complete P_t,
  * O_t read only after the item's last P V has completed; the next item's first P V only after all four warps read it,
  * no deadlock, every (tile, key block) processed exactly once, stream terminates on the sentinel.

    python tools/sim_attention_v2.py [n_seeds]

The waits use exactly the parity expressions of the kernel, so a barrier running two phases ahead of a waiter (parity
aliasing) shows up as a hazard or a deadlock here.
"""
import random
import sys

AQ, AK = 128, 64
KS, VS, TILES = 3, 2, 2


class Bar:
    def __init__(self, name, count):
        self.name, self.count = name, count
        self.pending, self.tx, self.completed = count, 0, 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.completed += 1
            self.pending = self.count

    def arrive(self, tx=0):
        assert self.pending > 0, f"{self.name}: more arrivals than the phase expects"
        self.tx += tx
        self.pending -= 1
        self._check()

    def complete_tx(self, n):
        self.tx -= n
        assert self.tx >= 0
        self._check()

    def passed(self, parity):   # mbarrier.try_wait.parity: true once the phase with this parity has completed
        return (self.completed & 1) != parity


def make_item(L, q0, window):
    nt = 2 if q0 + AQ < L else 1
    t_lo, t_hi = [], []
    for t in range(TILES):
        qa = q0 + t * AQ
        lo, hi = (0, L) if window < 0 else (max(0, qa - window), min(L, qa + AQ + window))
        t_lo.append(lo // AK)
        t_hi.append((hi + AK - 1) // AK)
    j_lo = min(t_lo[:nt])
    nb = max(t_hi[:nt]) - j_lo
    return dict(L=L, q0=q0, nt=nt, t_lo=t_lo, t_hi=t_hi, j_lo=j_lo, nb=nb)


def takes(it, t, jb):
    return t < it["nt"] and it["t_lo"][t] <= jb < it["t_hi"][t]


class Sim:
    def __init__(self, items, seed):
        self.rng = random.Random(seed)
        self.items = items + [None]                      # None = sentinel
        B = Bar
        self.q_full, self.q_empty = B("q_full", 1), B("q_empty", 1 + 4 * TILES)
        self.k_full = [B(f"k_full{i}", 1) for i in range(KS)]
        self.k_empty = [B(f"k_empty{i}", 1) for i in range(KS)]
        self.v_full = [B(f"v_full{i}", 1) for i in range(VS)]
        self.v_empty = [B(f"v_empty{i}", 1) for i in range(VS)]
        self.s_full = [B(f"s_full{t}", 1) for t in range(TILES)]
        self.s_empty = [B(f"s_empty{t}", 4) for t in range(TILES)]
        self.p_full = [B(f"p_full{t}", 4) for t in range(TILES)]
        self.pv_done = [B(f"pv_done{t}", 1) for t in range(TILES)]
        # resources
        self.info = [None, None]                         # published item index per slot
        self.info_readers = [set(), set()]               # who has read the current content
        self.q_item = None                               # item index whose Q tiles are in smem (None while loading)
        self.k_slot = [None] * KS                        # (item, block) held, or "loading"
        self.v_slot = [None] * VS
        self.reads = {"q": 0, "k": [0] * KS, "v": [0] * VS, "p": [0] * TILES}   # MMAs in flight reading the resource
        self.s_val = [None] * TILES                      # (item, block) of the S in TMEM
        self.s_loaded = [4] * TILES                      # softmax warps that have loaded the current S_t
        self.p_val = [[None] * 4 for _ in range(TILES)]  # (item, block) each warp wrote into P_t
        self.o_blocks = [[] for _ in range(TILES)]       # blocks accumulated into O_t for the current item
        self.o_item = [None] * TILES
        self.o_read = [4] * TILES                        # warps that have read the final O_t
        self.tc_queue = []                               # in-order tensor-core work: closures
        self.async_events = []                           # TMA completions: closures (any order)
        self.done_blocks = set()                         # (item, tile, block) whose output contribution completed
        self.out_written = set()                         # (item, tile, quarter)

    # ---- roles (generators yield ("wait", bar, parity) or None for a plain step) -------------------------------
    def producer(self):
        it_n, g = 0, 0
        pending_v = None
        for idx, it in enumerate(self.items[:-1]):
            yield ("wait", self.q_empty, (it_n & 1) ^ 1)
            slot = it_n & 1
            if self.info[slot] is not None:
                assert len(self.info_readers[slot]) == 1 + 4 * TILES, f"info slot {slot} republished before all reads"
            assert self.reads["q"] == 0, "Q overwritten while S products still read it"
            self.info[slot], self.info_readers[slot] = idx, set()
            self.q_item = "loading"
            self.q_full.arrive(tx=2)

            def q_landed(idx=idx):
                self.q_item = idx
                self.q_full.complete_tx(2)
            self.async_events.append(q_landed)
            it_n += 1
            yield None
            for i in range(it["nb"]):
                ks = g % KS
                yield ("wait", self.k_empty[ks], ((g // KS) & 1) ^ 1)
                assert self.reads["k"][ks] == 0, "K slot overwritten while an S product reads it"
                self.k_slot[ks] = "loading"
                self.k_full[ks].arrive(tx=1)

                def k_landed(ks=ks, key=(idx, it["j_lo"] + i)):
                    self.k_slot[ks] = key
                    self.k_full[ks].complete_tx(1)
                self.async_events.append(k_landed)
                yield None
                if g > 0:
                    yield from self._load_v(g - 1, pending_v)
                pending_v = (idx, it["j_lo"] + i)
                g += 1
        if g > 0:
            yield from self._load_v(g - 1, pending_v)
        yield ("wait", self.q_empty, (it_n & 1) ^ 1)
        slot = it_n & 1
        if self.info[slot] is not None:
            assert len(self.info_readers[slot]) == 1 + 4 * TILES, "sentinel overwrote an unread slot"
        self.info[slot], self.info_readers[slot] = len(self.items) - 1, set()
        self.q_full.arrive()

    def _load_v(self, gv, key):
        vs = gv % VS
        yield ("wait", self.v_empty[vs], ((gv // VS) & 1) ^ 1)
        assert self.reads["v"][vs] == 0, "V slot overwritten while a P V product reads it"
        self.v_slot[vs] = "loading"
        self.v_full[vs].arrive(tx=1)

        def v_landed(vs=vs, key=key):
            self.v_slot[vs] = key
            self.v_full[vs].complete_tx(1)
        self.async_events.append(v_landed)
        yield None

    def _read_info(self, it_n, who):
        idx = self.info[it_n & 1]
        assert idx == it_n, f"{who} read item {idx} where {it_n} was expected"
        self.info_readers[it_n & 1].add(who)
        return idx, self.items[idx]

    def mma(self):
        g, gs, gp = 0, [0] * TILES, [0] * TILES
        it_n = 0
        while True:
            yield ("wait", self.q_full, it_n & 1)
            idx, it = self._read_info(it_n, "mma")
            if it is None:
                return

            def issue_s(i):
                G = g + i
                ks, jb = G % KS, it["j_lo"] + i
                yield ("wait", self.k_full[ks], (G // KS) & 1)
                assert self.k_slot[ks] == (idx, jb), f"S reads K slot holding {self.k_slot[ks]}, wanted {(idx, jb)}"
                assert self.q_item == idx, "S reads Q of another item"
                for t in range(TILES):
                    if not takes(it, t, jb):
                        continue
                    yield ("wait", self.s_empty[t], (gs[t] & 1) ^ 1)
                    assert self.s_loaded[t] == 4, f"S_{t} overwritten before all four warps loaded it"
                    self.reads["q"] += 1
                    self.reads["k"][ks] += 1

                    def s_done(t=t, ks=ks, key=(idx, jb)):
                        self.s_val[t], self.s_loaded[t] = key, 0
                        self.reads["q"] -= 1
                        self.reads["k"][ks] -= 1
                    self.tc_queue.append(s_done)
                    self.tc_queue.append(self.s_full[t].arrive)
                    gs[t] += 1
                    yield None
                self.tc_queue.append(self.k_empty[ks].arrive)
                if i + 1 == it["nb"]:
                    self.tc_queue.append(self.q_empty.arrive)
                yield None

            yield from issue_s(0)
            for i in range(it["nb"]):
                if i + 1 < it["nb"]:
                    yield from issue_s(i + 1)
                G = g + i
                vs, jb = G % VS, it["j_lo"] + i
                yield ("wait", self.v_full[vs], (G // VS) & 1)
                assert self.v_slot[vs] == (idx, jb), f"P V reads V slot holding {self.v_slot[vs]}, wanted {(idx, jb)}"
                for t in range(TILES):
                    if not takes(it, t, jb):
                        continue
                    yield ("wait", self.p_full[t], gp[t] & 1)
                    assert all(v == (idx, jb) for v in self.p_val[t]), f"P V on an incomplete P_{t}: {self.p_val[t]}"
                    first = jb == it["t_lo"][t]
                    if first:
                        assert self.o_read[t] == 4, f"O_{t} overwritten before all four warps read the previous item's"
                    self.reads["v"][vs] += 1
                    self.reads["p"][t] += 1

                    def pv_done(t=t, vs=vs, first=first, key=(idx, jb)):
                        if first:
                            self.o_blocks[t], self.o_item[t], self.o_read[t] = [], key[0], 0
                        self.o_blocks[t].append(key[1])
                        self.reads["v"][vs] -= 1
                        self.reads["p"][t] -= 1
                        assert (key[0], t, key[1]) not in self.done_blocks
                        self.done_blocks.add((key[0], t, key[1]))
                    self.tc_queue.append(pv_done)
                    self.tc_queue.append(self.pv_done[t].arrive)
                    gp[t] += 1
                    yield None
                self.tc_queue.append(self.v_empty[vs].arrive)
                yield None
            g += it["nb"]
            it_n += 1

    def softmax(self, tile, quarter):
        kt, it_n = 0, 0
        who = f"soft{tile}{quarter}"
        while True:
            yield ("wait", self.q_full, it_n & 1)
            idx, it = self._read_info(it_n, who)
            self.q_empty.arrive()
            it_n += 1
            if it is None:
                return
            if tile >= it["nt"]:
                continue
            lo, hi = it["t_lo"][tile], it["t_hi"][tile]
            for i in range(hi - lo):
                G = kt + i
                yield ("wait", self.s_full[tile], G & 1)
                assert self.s_val[tile] == (idx, lo + i), f"{who} loads S {self.s_val[tile]}, wanted {(idx, lo + i)}"
                self.s_loaded[tile] += 1
                self.s_empty[tile].arrive()
                yield None
                if i > 0:
                    yield ("wait", self.pv_done[tile], (G - 1) & 1)
                assert self.reads["p"][tile] == 0, f"{who} overwrites P_{tile} while a P V product reads it"
                self.p_val[tile][quarter] = (idx, lo + i)
                yield None
                self.p_full[tile].arrive()
            G_last = kt + (hi - lo) - 1
            yield ("wait", self.pv_done[tile], G_last & 1)
            assert self.o_item[tile] == idx and self.o_blocks[tile] == list(range(lo, hi)), \
                f"{who} reads O_{tile} = item {self.o_item[tile]} blocks {self.o_blocks[tile]}, wanted {idx} {lo}..{hi}"
            self.o_read[tile] += 1
            self.out_written.add((idx, tile, quarter))
            yield None
            kt += hi - lo

    # ---- scheduler ------------------------------------------------------------------------------------------------
    def run(self, max_steps=2_000_000):
        roles = {"producer": self.producer(), "mma": self.mma()}
        for t in range(TILES):
            for q in range(4):
                roles[f"soft{t}{q}"] = self.softmax(t, q)
        waiting = {name: None for name in roles}          # pending ("wait", bar, parity)
        live = set(roles)
        # adversarial scheduling: a random subset of the roles is "slow" (scheduled only when nothing else can run, or
        # with a small probability) -- a warp that lags whole items behind is what exposes parity aliasing
        slow = {n for n in roles if self.rng.random() < 0.3}
        for _ in range(max_steps):
            runnable = [n for n in live if waiting[n] is None or waiting[n][1].passed(waiting[n][2])]
            fast = [n for n in runnable if n not in slow]
            if fast and self.rng.random() > 0.02:
                runnable = fast
            choices = [("role", n) for n in runnable]
            if self.tc_queue:
                choices.append(("tc", None))
            choices += [("async", k) for k in range(len(self.async_events))]
            if not choices:
                if live:
                    stuck = {n: (waiting[n][1].name, waiting[n][2], waiting[n][1].completed) for n in live}
                    raise AssertionError(f"deadlock: {stuck}")
                break
            kind, x = self.rng.choice(choices)
            self._running = kind
            if kind == "tc":
                self.tc_queue.pop(0)()
            elif kind == "async":
                self.async_events.pop(x)()
            else:
                waiting[x] = None
                try:
                    r = next(roles[x])
                    if r is not None:
                        waiting[x] = r
                except StopIteration:
                    live.discard(x)
        else:
            raise AssertionError("step limit")
        # completeness
        for idx, it in enumerate(self.items[:-1]):
            for t in range(it["nt"]):
                for jb in range(it["t_lo"][t], it["t_hi"][t]):
                    assert (idx, t, jb) in self.done_blocks, f"missing block {(idx, t, jb)}"
                for q in range(4):
                    assert (idx, t, q) in self.out_written, f"missing output {(idx, t, q)}"
        n_expected = sum(it["t_hi"][t] - it["t_lo"][t] for it in self.items[:-1] for t in range(it["nt"]))
        assert len(self.done_blocks) == n_expected


def random_items(rng, window):
    items = []
    for _ in range(rng.randint(0, 7)):
        L = rng.choice([1, 64, 100, 128, 129, 255, 256, 257, 300, 384, 385, 512, 513, 640, 700, 1033])
        for q0 in range(0, L, 2 * AQ):
            if rng.random() < 0.7:       # this CTA gets a random subset of the item heads
                items.append(make_item(L, q0, window))
    return items


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    total = 0
    for seed in range(n):
        rng = random.Random(10_000 + seed)
        window = rng.choice([-1, 64, 64, 0, 200])
        items = random_items(rng, window)
        Sim(items, seed).run()
        total += len(items)
    print(f"ok: {n} random schedules, {total} items, no hazard / deadlock")


if __name__ == "__main__":
    main()
