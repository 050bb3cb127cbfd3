"""Executable model of the mbarrier hand-offs of the two tcgen05 convolution kernels (csrc/conv_row.cu, csrc/conv_tc.cu).

Roles, as in the kernels: ONE loader thread issues the raw-slab loads in slab order (the loads complete in ANY order, as TMA loads do),
TWO transform groups take alternate slabs (raw stage -> operand stage), one or two MMA-issuing warps consume operand stages (two warps:
alternate tiles of `slabs_per_tile` slabs each).  Every wait is a PARITY wait: `test(P)` is true iff the barrier's current phase has the
other parity, i.e. "the phase with parity P has completed" - a waiter that polls while the barrier is still one use behind passes on
the phase before it.  The model runs random schedules and reports the first of

  * wrong slab   - a role consumed a stage that does not hold the slab it was waiting for (silent corruption in the kernel),
  * over-arrival - an arrive on a barrier whose phase has no arrival left (in the kernel: Warp Illegal Instruction at the loader's
                   next arrive.expect_tx - how the odd raw ring showed up),
  * deadlock     - nobody can move (in the kernel: a timed-out wait - how the two issuing warps showed up).

It is an argument about the PROTOCOL, kept next to the tests so that ring depths and role counts cannot be
changed without re-running it; it does not execute the kernels."""
import random


class Violation(Exception):
    pass


class Barrier:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _maybe_complete(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self, tx=0):
        if self.pending == 0:
            raise Violation("over-arrival")
        self.pending -= 1
        self.tx += tx
        self._maybe_complete()

    def complete_tx(self, tx):
        self.tx -= tx
        self._maybe_complete()

    def test(self, parity):
        return (self.phase & 1) != parity


def simulate(nr, na, slabs, slabs_per_tile=1, issuers=1, rng=None, max_steps=200000, per_warp_full=False):
    """One random schedule of `slabs` slabs through a raw ring of depth nr and an operand ring of depth na.  Raises Violation.
    per_warp_full: one full_a barrier set per issuing warp (the transform arrives on the set of the warp that owns the slab's tile, each
    warp counts its own uses of a stage) - the way to have two issuing warps without the parity hazard; not what the kernels do today."""
    rng = rng or random.Random(0)
    raw_full = [Barrier(1) for _ in range(nr)]
    raw_empty = [Barrier(1) for _ in range(nr)]      # one arrival per consuming GROUP (4 warps in the kernel)
    full_a = [[Barrier(1) for _ in range(na)] for _ in range(issuers if per_warp_full else 1)]
    uses = [[0] * na for _ in range(issuers)]        # per_warp_full: how often warp w has consumed stage s
    empty_a = [Barrier(1) for _ in range(na)]
    raw_content, a_content = [None] * nr, [None] * na
    in_flight = []                                   # issued loads that have not landed yet: (stage, slab)
    st = {"loader": 0, "xf": [0, 1], "xf_stage": [0, 0], "mma": [0] * issuers, "mma_tile": list(range(issuers))}
    done = {"loader": False, "xf": [False, False], "mma": [False] * issuers}
    # per-trial speeds: some roles (or the landing of loads) are much slower than others in some trials
    speed = {k: rng.choice([1, 1, 1, 5, 25]) for k in ("loader", "land", "xf0", "xf1", "mma0", "mma1")}
    tiles = (slabs + slabs_per_tile - 1) // slabs_per_tile

    def step_loader():
        q = st["loader"]
        if q >= slabs:
            done["loader"] = True
            return False
        s, k = q % nr, q // nr
        if not raw_empty[s].test((k & 1) ^ 1):
            return False
        raw_full[s].arrive(tx=1)                     # arrive.expect_tx, then the load is in flight
        in_flight.append((s, q))
        st["loader"] = q + 1
        return True

    def step_land():
        if not in_flight:
            return False
        s, q = in_flight.pop(rng.randrange(len(in_flight)))   # loads complete out of order
        raw_content[s] = q
        raw_full[s].complete_tx(1)
        return True

    def step_xf(g):
        q = st["xf"][g]
        if q >= slabs:
            done["xf"][g] = True
            return False
        s, sa = q % nr, q % na
        if st["xf_stage"][g] == 0:                   # wait raw_full, then (a separate poll, possibly much later) empty_a
            if not raw_full[s].test((q // nr) & 1):
                return False
            st["xf_stage"][g] = 1
            return True
        if not empty_a[sa].test(((q // na) & 1) ^ 1):
            return False
        if raw_content[s] != q:
            raise Violation("wrong slab in raw stage %d: transform group %d wanted %d, found %s" % (s, g, q, raw_content[s]))
        a_content[sa] = q
        full_a[(q // slabs_per_tile) % issuers if per_warp_full else 0][sa].arrive()
        raw_empty[s].arrive()
        st["xf"][g], st["xf_stage"][g] = q + 2, 0
        return True

    def step_mma(w):
        t = st["mma_tile"][w]
        if t >= tiles:
            done["mma"][w] = True
            return False
        q = t * slabs_per_tile + st["mma"][w]
        if q >= slabs:
            done["mma"][w] = True
            return False
        sa = q % na
        if per_warp_full:
            if not full_a[w][sa].test(uses[w][sa] & 1):
                return False
            uses[w][sa] += 1
        elif not full_a[0][sa].test((q // na) & 1):
            return False
        if a_content[sa] != q:
            raise Violation("wrong slab in operand stage %d: issuing warp %d wanted %d, found %s" % (sa, w, q, a_content[sa]))
        empty_a[sa].arrive()                         # tcgen05.commit
        st["mma"][w] += 1
        if st["mma"][w] == slabs_per_tile:
            st["mma"][w], st["mma_tile"][w] = 0, t + issuers
        return True

    agents = [("loader", step_loader), ("land", step_land), ("xf0", lambda: step_xf(0)), ("xf1", lambda: step_xf(1))]
    agents += [("mma%d" % w, (lambda w=w: step_mma(w))) for w in range(issuers)]
    idle = 0
    for _ in range(max_steps):
        if done["loader"] and all(done["xf"]) and all(done["mma"]) and not in_flight:
            return
        name, fn = agents[rng.randrange(len(agents))]
        if rng.randrange(speed[name]) != 0:          # slow roles skip most of their turns
            continue
        if fn():
            idle = 0
        else:
            idle += 1
            if idle > 4000 and not any(f() for _, f in agents):
                if done["loader"] and all(done["xf"]) and all(done["mma"]) and not in_flight:
                    return
                raise Violation("deadlock")
    raise Violation("no progress within %d steps" % max_steps)


def first_violation(nr, na, slabs=48, slabs_per_tile=1, issuers=1, trials=400, seed=0, per_warp_full=False):
    """None, or the message of the first violation found over `trials` random schedules."""
    for i in range(trials):
        try:
            simulate(nr, na, slabs, slabs_per_tile, issuers, random.Random(seed * 100003 + i), per_warp_full=per_warp_full)
        except Violation as v:
            return "trial %d: %s" % (i, v)
    return None


def simulate_acc(ns, rows, rng=None, max_steps=200000):
    """Accumulator hand-off of conv_row_kernel: the MMA warp fills one TMEM slot per INPUT row (ring of ns slots), two epilogue groups
    own the even / odd OUTPUT rows; output row y reads the slots of input rows y - 1, y, y + 1 and waits only for the last of them.
    A slot is handed back through a barrier that counts 3 arrivals per use (one per reading output row); the first / last row of
    the run arrives for the rows that never come.  Raises Violation."""
    rng = rng or random.Random(0)
    acc_full = [Barrier(1) for _ in range(ns)]
    acc_empty = [Barrier(3) for _ in range(ns)]
    content = [None] * ns
    st = {"mma": 0, "epi": [0, 1]}
    speed = {k: rng.choice([1, 1, 4, 20]) for k in ("mma", "epi0", "epi1")}

    def step_mma():
        j = st["mma"]
        if j >= rows:
            return False
        s = j % ns
        if not acc_empty[s].test(((j // ns) & 1) ^ 1):
            return False
        content[s] = j
        acc_full[s].arrive()                         # tcgen05.commit
        st["mma"] = j + 1
        return True

    def step_epi(e):
        y = st["epi"][e]
        if y >= rows:
            return False
        last = min(y + 1, rows - 1)
        if not acc_full[last % ns].test((last // ns) & 1):
            return False
        first, lastrow = y == 0, y == rows - 1
        for j, n in ((y - 1, 3 if first else 1), (y, 1 + first + lastrow), (y + 1, 3 if lastrow else 1)):
            if 0 <= j < rows:
                if content[j % ns] != j:
                    raise Violation("wrong accumulator in slot %d: output row %d wanted input row %d, found %s" % (j % ns, y, j, content[j % ns]))
                for _ in range(n):
                    acc_empty[j % ns].arrive()
        st["epi"][e] = y + 2
        return True

    agents = [("mma", step_mma), ("epi0", lambda: step_epi(0)), ("epi1", lambda: step_epi(1))]
    idle = 0
    for _ in range(max_steps):
        if st["mma"] >= rows and st["epi"][0] >= rows and st["epi"][1] >= rows:
            return
        name, fn = agents[rng.randrange(3)]
        if rng.randrange(speed[name]) != 0:
            continue
        if fn():
            idle = 0
        else:
            idle += 1
            if idle > 2000 and not any(f() for _, f in agents):
                raise Violation("deadlock")
    raise Violation("no progress within %d steps" % max_steps)


def first_acc_violation(ns, rows=40, trials=300, seed=0):
    for i in range(trials):
        try:
            simulate_acc(ns, rows, random.Random(seed * 100003 + i))
        except Violation as v:
            return "trial %d: %s" % (i, v)
    return None


def explore(nr, na, slabs, slabs_per_tile=1, issuers=1, max_states=2000000):
    """EXHAUSTIVE search over every interleaving of the roles of `simulate` (shared full_a set) for a small number of slabs.
    Returns (None, states visited) or (violation message, states visited)."""
    tiles = (slabs + slabs_per_tile - 1) // slabs_per_tile
    # a barrier is (pending, tx, phase); count is 1 everywhere in this part of the protocol
    def arrive(b, tx=0):
        pending, t, ph = b
        if pending == 0:
            raise Violation("over-arrival")
        pending, t = pending - 1, t + tx
        return (1, 0, ph + 1) if pending == 0 and t == 0 else (pending, t, ph)

    def complete(b):
        pending, t, ph = b
        t -= 1
        return (1, 0, ph + 1) if pending == 0 and t == 0 else (pending, t, ph)

    def test(b, parity):
        return (b[2] & 1) != parity

    def put(tup, i, v):
        return tup[:i] + (v,) + tup[i + 1:]

    fresh = (1, 0, 0)
    # state: loader, in_flight (sorted tuple of (stage, slab)), xf (q0, stage0, q1, stage1), mma ((tile, idx) per warp),
    #        raw_full, raw_empty, full_a, empty_a, raw_content, a_content
    init = (0, (), (0, 0, 1, 0), tuple((w, 0) for w in range(issuers)), (fresh,) * nr, (fresh,) * nr, (fresh,) * na, (fresh,) * na,
            (None,) * nr, (None,) * na)
    seen, stack = {init}, [init]
    while stack:
        state = stack.pop()
        ld, fl, xf, mma, rf, re_, fa, ea, rc, ac = state
        succ = []
        try:
            if ld < slabs:                                             # loader
                s, k = ld % nr, ld // nr
                if test(re_[s], (k & 1) ^ 1):
                    succ.append((ld + 1, tuple(sorted(fl + ((s, ld),))), xf, mma, put(rf, s, arrive(rf[s], 1)), re_, fa, ea, rc, ac))
            for i, (s, q) in enumerate(fl):                            # any load in flight may land next
                succ.append((ld, fl[:i] + fl[i + 1:], xf, mma, put(rf, s, complete(rf[s])), re_, fa, ea, put(rc, s, q), ac))
            for g in (0, 1):                                           # transform groups
                q, stg = xf[2 * g], xf[2 * g + 1]
                if q >= slabs:
                    continue
                s, sa = q % nr, q % na
                if stg == 0:
                    if test(rf[s], (q // nr) & 1):
                        succ.append((ld, fl, put(xf, 2 * g + 1, 1), mma, rf, re_, fa, ea, rc, ac))
                elif test(ea[sa], ((q // na) & 1) ^ 1):
                    if rc[s] != q:
                        raise Violation("wrong slab in raw stage %d: transform group %d wanted %d, found %s" % (s, g, q, rc[s]))
                    nxf = put(put(xf, 2 * g, q + 2), 2 * g + 1, 0)
                    succ.append((ld, fl, nxf, mma, rf, put(re_, s, arrive(re_[s])), put(fa, sa, arrive(fa[sa])), ea, rc, put(ac, sa, q)))
            for w in range(issuers):                                   # issuing warps
                t, i = mma[w]
                q = t * slabs_per_tile + i
                if t >= tiles or q >= slabs:
                    continue
                sa = q % na
                if test(fa[sa], (q // na) & 1):
                    if ac[sa] != q:
                        raise Violation("wrong slab in operand stage %d: issuing warp %d wanted %d, found %s" % (sa, w, q, ac[sa]))
                    nm = (t, i + 1) if i + 1 < slabs_per_tile else (t + issuers, 0)
                    succ.append((ld, fl, xf, put(mma, w, nm), rf, re_, fa, put(ea, sa, arrive(ea[sa])), rc, ac))
        except Violation as v:
            return str(v), len(seen)
        finished = ld >= slabs and not fl and xf[0] >= slabs and xf[2] >= slabs and all(
            t >= tiles or t * slabs_per_tile + i >= slabs for t, i in mma)
        if not succ and not finished:
            return "deadlock", len(seen)
        for n in succ:
            if n not in seen:
                if len(seen) >= max_states:
                    raise RuntimeError("state space larger than %d" % max_states)
                seen.add(n)
                stack.append(n)
    return None, len(seen)
