"""Randomised interleaving check of the mbarrier protocol of the fused bottleneck-exit kernel (one CTA pair).

Every role of the kernel (TMA producers, the two MMA-issuing threads, the epilogue warps of both CTAs) is a coroutine that
blocks on `wait(barrier, parity)`; arrivals that the hardware delivers asynchronously (tcgen05.commit after the MMAs retire, TMA
complete_tx) go through per-issuer FIFO queues that a scheduler drains at random times.  A run that ends with blocked roles and
empty queues is a deadlock; a wait whose barrier ran two phases ahead (parity aliasing) shows up the same way.

    python tools/protocol_sim.py [variant] [runs]      variant: v3 (library kernel) | v5 (tools/probe variant)
"""
import random
import sys


class Bar:
    def __init__(self, name, count):
        self.name, self.init, self.pending, self.tx, self.k = name, count, count, 0, 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.k += 1
            self.pending = self.init

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "too many arrivals on %s" % self.name
        self._check()

    def expect(self, n):
        self.tx += n
        self.arrive()

    def complete(self, n):
        self.tx -= n
        self._check()

    def passed(self, parity):
        return (self.k & 1) != parity


def simulate(variant, seed, tiles=3, NCH=16, kRing=3, kAcc=4, store_y=True, skew=None, p_deliver=0.25):
    """skew: None | a predicate on the role name; roles it selects run 20x less often (a CTA, a warp, a producer that lags).
    p_deliver: probability of delivering a deferred arrival when roles are runnable (0 = as late as possible)."""
    rnd = random.Random(seed)
    groups = 3 if variant == "v5" else 2
    L = {}   # leader barriers
    C = [{}, {}]   # per-CTA barriers

    def mk(d, name, count, n=None):
        if n is None:
            d[name] = Bar(name, count)
        else:
            d[name] = [Bar("%s[%d]" % (name, i), count) for i in range(n)]
    mk(L, "h_full", 1); mk(L, "b1_full", 1, kRing); mk(L, "b2_full", 1, kRing)
    mk(L, "acc1_empty", 2 * 4, kAcc); mk(L, "a2_full", 2 * 4, kAcc); mk(L, "acc2_empty", 2 * 4 * groups)
    for r in (0, 1):
        mk(C[r], "h_empty", 1); mk(C[r], "b1_empty", 1, kRing); mk(C[r], "b2_empty", 1, kRing)
        mk(C[r], "acc1_full", 1, kAcc); mk(C[r], "a2_empty", 1, kAcc); mk(C[r], "acc2_full", 1)
        mk(C[r], "xa_full", 1, kAcc); mk(C[r], "a2_free", 4, kAcc)
    queues = {}   # issuer -> FIFO of deferred arrivals (in-order per issuer)

    def defer(issuer, fn):
        queues.setdefault(issuer, []).append(fn)

    def both(name, idx=None):
        def f():
            for r in (0, 1):
                b = C[r][name] if idx is None else C[r][name][idx]
                b.arrive()
        return f

    W = ("wait",)

    def prod_h_w3(r):
        cc = 0
        for it in range(tiles):
            yield (W, C[r]["h_empty"], (it & 1) ^ 1)
            if r == 0:
                L["h_full"].expect(2)
            defer(("tma", r, "h"), lambda: L["h_full"].complete(1))
            for c in range(NCH):
                s = cc % kRing
                yield (W, C[r]["b1_empty"][s], ((cc // kRing) & 1) ^ 1)
                if r == 0:
                    L["b1_full"][s].expect(2)
                defer(("tma", r, "w3"), (lambda s=s: L["b1_full"][s].complete(1)))
                cc += 1

    def prod_w1(r):
        cc = 0
        for it in range(tiles):
            for c in range(NCH):
                s = cc % kRing
                yield (W, C[r]["b2_empty"][s], ((cc // kRing) & 1) ^ 1)
                if r == 0:
                    L["b2_full"][s].expect(2)
                defer(("tma", r, "w1"), (lambda s=s: L["b2_full"][s].complete(1)))
                cc += 1

    def prod_x(r):
        cc = 0
        for it in range(tiles):
            for c in range(NCH):
                b, ph = cc % kAcc, (cc // kAcc) & 1
                yield (W, C[r]["a2_empty"][b], ph ^ 1)
                yield (W, C[r]["a2_free"][b], ph ^ 1)
                C[r]["xa_full"][b].expect(1)
                defer(("tma", r, "x"), (lambda b=b: C[r]["xa_full"][b].complete(1)))
                cc += 1

    def gemm1():
        cc = 0
        for it in range(tiles):
            yield (W, L["h_full"], it & 1)
            for c in range(NCH):
                s, b = cc % kRing, cc % kAcc
                yield (W, L["b1_full"][s], (cc // kRing) & 1)
                yield (W, L["acc1_empty"][b], ((cc // kAcc) & 1) ^ 1)
                defer("mma1", both("b1_empty", s))
                defer("mma1", both("acc1_full", b))
                cc += 1
            defer("mma1", both("h_empty"))

    def gemm2():
        cc = 0
        for it in range(tiles):
            for c in range(NCH):
                s, b = cc % kRing, cc % kAcc
                yield (W, L["b2_full"][s], (cc // kRing) & 1)
                yield (W, L["a2_full"][b], (cc // kAcc) & 1)
                if c == 0:
                    yield (W, L["acc2_empty"], (it & 1) ^ 1)
                defer("mma2", both("b2_empty", s))
                defer("mma2", both("a2_empty", b))
                cc += 1
            defer("mma2", both("acc2_full"))

    def epi(r, g, q):
        for it in range(tiles):
            cc0 = it * NCH
            pending = [-1]

            def release():
                if pending[0] >= 0:
                    C[r]["a2_free"][pending[0]].arrive()
                    pending[0] = -1
            for c in range(g, NCH, groups):
                cc = cc0 + c
                b, ph = cc % kAcc, (cc // kAcc) & 1
                if variant == "v5":
                    release()
                yield (W, C[r]["acc1_full"][b], ph)
                L["acc1_empty"][b].arrive()
                if variant == "v5":
                    yield (W, C[r]["xa_full"][b], ph)
                else:
                    yield (W, C[r]["a2_empty"][b], ph ^ 1)
                yield ("yield",)                      # the math phase: lets other roles run in between
                L["a2_full"][b].arrive()
                if variant == "v5":
                    if store_y:
                        pending[0] = b
                    else:
                        C[r]["a2_free"][b].arrive()
            if variant == "v5":
                release()
            yield (W, C[r]["acc2_full"], it & 1)
            L["acc2_empty"].arrive()

    roles = {}
    for r in (0, 1):
        roles[("h_w3", r)] = prod_h_w3(r)
        roles[("w1", r)] = prod_w1(r)
        if variant == "v5":
            roles[("x", r)] = prod_x(r)
        for g in range(groups):
            for q in range(4):
                roles[("epi", r, g, q)] = epi(r, g, q)
    roles["gemm1"] = gemm1()
    roles["gemm2"] = gemm2()
    blocked = {}          # role -> (bar, parity)
    for name in list(roles):
        blocked[name] = None
    live = set(roles)
    steps = 0
    while live:
        runnable = [n for n in live if blocked[n] is None or blocked[n][0].passed(blocked[n][1])]
        drains = [k for k, q in queues.items() if q]
        if not runnable and not drains:
            return ("deadlock", {n: (blocked[n][0].name, blocked[n][1], blocked[n][0].k) for n in live})
        # random choice between running a role and delivering a deferred arrival (deliveries are often late)
        if drains and (not runnable or rnd.random() < p_deliver):
            k = rnd.choice(drains)
            queues[k].pop(0)()
            continue
        if skew is not None and len(runnable) > 1:
            fast = [n for n in runnable if not skew(n)]
            if fast and rnd.random() < 0.95:
                runnable = fast
        n = rnd.choice(runnable)
        blocked[n] = None
        try:
            ev = next(roles[n])
        except StopIteration:
            live.discard(n)
            continue
        if ev[0] is W:
            blocked[n] = (ev[1], ev[2])
        steps += 1
        if steps > 2_000_000:
            return ("runaway", None)
    return ("ok", None)


if __name__ == "__main__":
    variant = sys.argv[1] if len(sys.argv) > 1 else "v5"
    runs = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    SKEWS = {
        "none": None,
        "cta1 slow": lambda n: isinstance(n, tuple) and len(n) > 1 and n[1] == 1,
        "cta0 slow": lambda n: isinstance(n, tuple) and len(n) > 1 and n[1] == 0,
        "one epilogue warp slow": lambda n: n == ("epi", 1, 0, 2),
        "producers slow": lambda n: isinstance(n, tuple) and n[0] in ("h_w3", "w1", "x"),
        "mma threads slow": lambda n: n in ("gemm1", "gemm2"),
        "epilogue slow": lambda n: isinstance(n, tuple) and n[0] == "epi",
    }
    total = bad = 0
    for sk_name, sk in SKEWS.items():
        for p_deliver in (0.25, 0.0, 0.9):
            for store_y in (True, False):
                for seed in range(runs):
                    res, info = simulate(variant, seed, store_y=store_y, skew=sk, p_deliver=p_deliver)
                    total += 1
                    if res != "ok":
                        bad += 1
                        if bad <= 3:
                            print(variant, "skew", sk_name, "p_deliver", p_deliver, "store_y", store_y, "seed", seed, res)
                            for k, v in sorted(info.items(), key=str)[:40]:
                                print("   ", k, "waits on", v)
    print(variant, "runs", total, "failures", bad)
