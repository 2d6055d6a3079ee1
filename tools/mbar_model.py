"""A small executable model of the mbarrier / tcgen05 pipelines of the attention kernels (test infrastructure).

Why: the one dead-lock of round 1 was a PARITY ALIASING bug — a consumer group arrived twice on a shared barrier before
its peers had arrived once, the phase completed with the wrong membership, and a parity wait that was two phases behind
blocked for ever.  Such bugs do not show in a unit test with warm caches; they show under a different interleaving.
This model runs the protocols of `csrc/attention.cu` / `csrc/attention_r2.cu` under thousands of random interleavings
and checks, at every wait, that the waiter is AT MOST ONE PHASE BEHIND the barrier (the condition under which a parity
wait means what the code thinks it means), plus the data hazards the barriers exist for.

Semantics modelled (PTX ISA, `mbarrier.try_wait.parity`, `tcgen05.commit`):
  * a barrier has an arrival count; the phase completes when `count` arrivals have been made; `completed` phases so far;
  * `wait(parity p)` succeeds iff the current phase parity differs from p, i.e. iff `(completed & 1) != p`;
  * tcgen05.mma instructions of one issuing thread execute in issue order; `tcgen05.commit` arrives on a barrier when
    every MMA issued before it has completed — modelled as a FIFO drained by its own agent;
  * TMA loads complete asynchronously (their own agent performs the barrier's transaction arrival).
Agents are Python generators yielding actions; a seeded scheduler picks any runnable agent for the next action.
"""
from __future__ import annotations

import random
from collections import deque


class ProtocolError(AssertionError):
    pass


class Barrier:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.completed = name, count, 0, 0

    def arrive(self):
        self.pending += 1
        if self.pending == self.count:
            self.pending = 0
            self.completed += 1

    def test(self, parity, phase_index):
        """Parity wait for the completion of phase `phase_index` (0-based; -1 = "the phase before the first", the
        `ph ^ 1` idiom for buffers that start empty).  Returns what the hardware returns, `(completed & 1) != parity`,
        and raises when that answer would not mean "phase_index has completed": the waiter must be inside the window
        completed in {phase_index, phase_index + 1}."""
        if parity != (phase_index & 1):
            raise ProtocolError(f"{self.name}: parity expression {parity} does not match phase {phase_index}")
        if self.completed > phase_index + 1:
            raise ProtocolError(f"{self.name}: waiter for phase {phase_index} was lapped (completed {self.completed}): "
                                f"the parity test now blocks although the phase is long over")
        if self.completed < phase_index:
            raise ProtocolError(f"{self.name}: wait for phase {phase_index} would return early (completed {self.completed})")
        return (self.completed & 1) != parity


class Sim:
    """Scheduler.  An agent yields ('wait', barrier, parity, phase_index) | ('arrive', barrier) |
    ('do', callable) | ('push', item) (to the tensor pipe) | ('sync', key, n) (named CTA barrier of n agents)."""

    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.agents = {}
        self.weight = {}
        self.pipe = deque()
        self.sync_wait = {}
        self.steps = 0

    def add(self, name, gen):
        self.agents[name] = gen
        # skewed speeds: an agent may be 100x slower or faster than its peers for the whole run (a delayed MMA warp, a
        # slow TMA, one straggling softmax warp are exactly the interleavings a warm unit test never produces)
        self.weight[name] = self.rng.choice((0.02, 0.2, 1.0, 1.0, 5.0))

    def pipe_agent(self):
        """Drains the tensor pipe in issue order: ('mma', fn) executes fn; ('commit', barrier) arrives."""
        while True:
            if not self.pipe:
                yield ("idle",)
                continue
            kind, x = self.pipe.popleft()
            if kind == "mma":
                yield ("do", x)
            else:
                yield ("arrive", x)

    def run(self, max_steps=2_000_000):
        self.add("tensor_pipe", self.pipe_agent())
        done = set()
        pending_action = {}
        while True:
            runnable = []
            for name in self.agents:
                if name in done:
                    continue
                act = pending_action.get(name)
                if act is None:
                    try:
                        act = next(self.agents[name])
                    except StopIteration:
                        done.add(name)
                        continue
                    pending_action[name] = act
                if act[0] == "wait":
                    _, bar, parity, idx = act
                    if bar.test(parity, idx):
                        runnable.append(name)
                elif act[0] == "idle":
                    if self.pipe:
                        pending_action[name] = None
                        runnable.append(name)
                elif act[0] == "sync":
                    _, key, n = act
                    waiting = self.sync_wait.setdefault(key, set())
                    waiting.add(name)
                    if len(waiting) == n:
                        runnable.append(name)
                else:
                    runnable.append(name)
            workers = [n for n in self.agents if n not in done and n != "tensor_pipe"]
            if not workers and not self.pipe:
                return
            if not runnable:
                state = {n: pending_action.get(n) for n in self.agents if n not in done}
                raise ProtocolError(f"deadlock: {[(n, a[:1] + tuple(getattr(x, 'name', x) for x in a[1:])) for n, a in state.items() if a]}")
            name = self.rng.choices(runnable, weights=[self.weight[n] for n in runnable])[0]
            act = pending_action[name]
            if act is None:
                continue
            if act[0] == "wait":
                pass
            elif act[0] == "arrive":
                act[1].arrive()
            elif act[0] == "do":
                act[1]()
            elif act[0] == "push":
                self.pipe.append(act[1])
            elif act[0] == "sync":
                _, key, n = act
                # release everybody waiting on this named barrier at once
                for other in list(self.sync_wait[key]):
                    pending_action[other] = None
                self.sync_wait[key] = set()
            pending_action[name] = None
            self.steps += 1
            if self.steps > max_steps:
                raise ProtocolError("step limit (livelock?)")


# ---------------------------------------------------------------------------------------------------------------------
# forward: attention_r2.cu attn_fwd_nt_kernel (lazy P V waits) and its single-p_full variant (the bug it avoids)
# ---------------------------------------------------------------------------------------------------------------------
def run_fwd(seed, n_kv=7, n_soft=3, lazy=True, double_p_full=True, rescale_prob=0.3, sync_group=None):
    """`sync_group`: size of the groups that exchange the row maximum through a named barrier (None = all softmax agents,
    as in attention.cu; attention_r2.cu synchronises only the NT warps that share a lane quarter)."""
    sim = Sim(seed)
    B = lambda n, c: Barrier(n, c)  # noqa: E731
    k_full, k_empty = [B(f"k_full{s}", 1) for s in (0, 1)], [B(f"k_empty{s}", 1) for s in (0, 1)]
    v_full, v_empty = [B(f"v_full{s}", 1) for s in (0, 1)], [B(f"v_empty{s}", 1) for s in (0, 1)]
    s_full, s_empty = [B(f"s_full{s}", 1) for s in (0, 1)], [B(f"s_empty{s}", n_soft) for s in (0, 1)]
    p_full = [B(f"p_full{s}", n_soft) for s in (0, 1)] if double_p_full else [B("p_full", n_soft)] * 2
    pv_done = [B(f"pv_done{s}", 1) for s in (0, 1)] if lazy else [B("pv_done", 1)] * 2
    st = dict(S_tile=[None, None], S_read=[set(), set()], P_tile=[None, None], P_written=[set(), set()], pv_exec=-1, s_exec=-1,
              k_tile=[None, None], v_tile=[None, None])
    rng = random.Random(seed * 7919 + 1)

    def tma():
        for j in range(n_kv):
            s, ph = j & 1, (j >> 1) & 1
            yield ("wait", k_empty[s], ph ^ 1, (j >> 1) - 1)
            yield ("do", lambda j=j, s=s: st["k_tile"].__setitem__(s, j))
            yield ("arrive", k_full[s])
            yield ("wait", v_empty[s], ph ^ 1, (j >> 1) - 1)
            yield ("do", lambda j=j, s=s: st["v_tile"].__setitem__(s, j))
            yield ("arrive", v_full[s])

    def mma():
        def exec_S(j):
            s = j & 1
            if st["k_tile"][s] != j:
                raise ProtocolError(f"S({j}) reads K buffer holding {st['k_tile'][s]}")
            if st["S_tile"][s] is not None and len(st["S_read"][s]) != n_soft:
                raise ProtocolError(f"S({j}) overwrites S({st['S_tile'][s]}) before every softmax agent loaded it")
            st["S_tile"][s], st["S_read"][s], st["s_exec"] = j, set(), j

        def exec_PV(j):
            s = j & 1
            if st["v_tile"][s] != j:
                raise ProtocolError(f"PV({j}) reads V buffer holding {st['v_tile'][s]}")
            if st["P_tile"][s] != j or len(st["P_written"][s]) != n_soft:
                raise ProtocolError(f"PV({j}) reads P buffer {st['P_tile'][s]} written by {len(st['P_written'][s])}/{n_soft}")
            st["pv_exec"] = j

        def issue_S(j):
            s, ph = j & 1, (j >> 1) & 1
            yield ("wait", k_full[s], ph, j >> 1)
            yield ("wait", s_empty[s], ph ^ 1, (j >> 1) - 1)
            yield ("push", ("mma", lambda j=j: exec_S(j)))
            yield ("push", ("commit", k_empty[s]))
            yield ("push", ("commit", s_full[s]))

        def issue_PV(j):
            s, ph = j & 1, (j >> 1) & 1
            yield ("wait", v_full[s], ph, j >> 1)
            if double_p_full:
                yield ("wait", p_full[s], ph, j >> 1)
            else:
                yield ("wait", p_full[0], j & 1, j)
            yield ("push", ("mma", lambda j=j: exec_PV(j)))
            yield ("push", ("commit", v_empty[s]))
            yield ("push", ("commit", pv_done[s]))

        yield from issue_S(0)
        for j in range(n_kv):
            if j + 1 < n_kv:
                yield from issue_S(j + 1)
            yield from issue_PV(j)

    def softmax(w):
        for j in range(n_kv):
            s, ph = j & 1, (j >> 1) & 1
            yield ("wait", s_full[s], ph, j >> 1)

            def load_S(j=j, s=s):
                if st["S_tile"][s] != j:
                    raise ProtocolError(f"softmax {w} loads S buffer holding {st['S_tile'][s]} instead of {j}")
                st["S_read"][s].add(w)
            yield ("do", load_S)
            yield ("arrive", s_empty[s])
            g = sync_group or n_soft
            yield ("sync", ("max", j, w // g), min(g, n_soft - (w // g) * g))  # bar.sync: row-max exchange
            if not lazy and j > 0:
                yield ("wait", pv_done[0], (j - 1) & 1, j - 1)
            if j > 0 and lazy and rng.random() < rescale_prob:
                yield ("wait", pv_done[s ^ 1], ((j - 1) >> 1) & 1, (j - 1) >> 1)

                def rescale(j=j):
                    if st["pv_exec"] != j - 1:
                        raise ProtocolError(f"softmax {w} rescales O at tile {j} while PV executed up to {st['pv_exec']}")
                yield ("do", rescale)
            if lazy and j >= 2:
                yield ("wait", pv_done[s], ((j >> 1) - 1) & 1, (j >> 1) - 1)

            def write_P(j=j, s=s):
                if st["pv_exec"] < j - 2:
                    raise ProtocolError(f"softmax {w} overwrites P({j - 2}) before PV({j - 2}) finished")
                if st["P_tile"][s] != j:
                    st["P_tile"][s], st["P_written"][s] = j, set()
                st["P_written"][s].add(w)
            yield ("do", write_P)
            yield ("arrive", p_full[s] if double_p_full else p_full[0])
        last = n_kv - 1
        if lazy:
            yield ("wait", pv_done[last & 1], (last >> 1) & 1, last >> 1)
        else:
            yield ("wait", pv_done[0], last & 1, last)

        def epilogue():
            if st["pv_exec"] != last:
                raise ProtocolError("epilogue before the last PV")
        yield ("do", epilogue)

    sim.add("tma", tma())
    sim.add("mma", mma())
    for w in range(n_soft):
        sim.add(f"soft{w}", softmax(w))
    sim.run()
    return sim.steps


# ---------------------------------------------------------------------------------------------------------------------
# backward: attn_bwd_kernel / attn_bwd_r2_kernel (two softmax groups, each owning TMEM buffer X[g] and pb_full[g]) and the
# round-1 bug (one shared pb_full for both groups)
# ---------------------------------------------------------------------------------------------------------------------
def run_bwd(seed, n_t=9, group_size=2, stages=3, shared_pb_full=False):
    sim = Sim(seed)
    t_full = [Barrier(f"t_full{s}", 1) for s in range(stages)]
    t_empty = [Barrier(f"t_empty{s}", 1) for s in range(stages)]
    x_full = [Barrier(f"x_full{g}", 1) for g in (0, 1)]
    pb_full = [Barrier("pb_full", 2 * group_size)] * 2 if shared_pb_full else [Barrier(f"pb_full{g}", group_size) for g in (0, 1)]
    done_bar = Barrier("done", 1)
    st = dict(T=[None] * stages, X=[None, None], X_kind=["", ""], P_written=[set(), set()], b_exec=-1)

    def tma():
        for i in range(n_t):
            s = i % stages
            yield ("wait", t_empty[s], ((i // stages) & 1) ^ 1, i // stages - 1)
            yield ("do", lambda i=i, s=s: st["T"].__setitem__(s, i))
            yield ("arrive", t_full[s])

    def mma():
        def exec_A(i):
            g = i & 1
            if st["T"][i % stages] != i:
                raise ProtocolError(f"A({i}) reads stage holding {st['T'][i % stages]}")
            if st["X"][g] is not None and st["b_exec"] < st["X"][g]:
                raise ProtocolError(f"A({i}) overwrites X[{g}] before B({st['X'][g]}) read P/dS from it")
            st["X"][g], st["X_kind"][g], st["P_written"][g] = i, "S", set()

        def exec_B(i):
            g = i & 1
            if st["T"][i % stages] != i:
                raise ProtocolError(f"B({i}) reads stage holding {st['T'][i % stages]}")
            if st["X"][g] != i or len(st["P_written"][g]) != group_size:
                raise ProtocolError(f"B({i}) reads X[{g}] = tile {st['X'][g]} with {len(st['P_written'][g])}/{group_size} rows of P")
            st["b_exec"] = i

        def issue_A(i):
            s = i % stages
            yield ("wait", t_full[s], (i // stages) & 1, i // stages)
            yield ("push", ("mma", lambda i=i: exec_A(i)))
            yield ("push", ("commit", x_full[i & 1]))

        def issue_B(i):
            if shared_pb_full:
                yield ("wait", pb_full[0], i & 1, i)
            else:
                yield ("wait", pb_full[i & 1], (i >> 1) & 1, i >> 1)
            yield ("push", ("mma", lambda i=i: exec_B(i)))
            yield ("push", ("commit", t_empty[i % stages]))

        yield from issue_A(0)
        for i in range(n_t):
            if i + 1 < n_t:
                yield from issue_A(i + 1)
            yield from issue_B(i)
        yield ("push", ("commit", done_bar))

    def softmax(g, w):
        for i in range(g, n_t, 2):
            yield ("wait", x_full[g], (i >> 1) & 1, i >> 1)

            def work(i=i):
                if st["X"][g] != i:
                    raise ProtocolError(f"group {g} reads X holding tile {st['X'][g]} instead of {i}")
                st["P_written"][g].add(w)
            yield ("do", work)
            yield ("arrive", pb_full[g])
        yield ("wait", done_bar, 0, 0)

    sim.add("tma", tma())
    sim.add("mma", mma())
    for g in (0, 1):
        for w in range(group_size):
            sim.add(f"soft{g}_{w}", softmax(g, w))
    sim.run()
    return sim.steps


# ---------------------------------------------------------------------------------------------------------------------
# backward dQ pass with EARLY issue of the next S/dP MMA (attention_r2.cu attn_bwd_q2_kernel): dS goes to its own TMEM
# region D[g], so A(i+2) may overwrite X[g] as soon as softmax(i) has LOADED S/dP (x_free) instead of after B(i)
# ---------------------------------------------------------------------------------------------------------------------
def run_bwd_q2(seed, n_t=9, group_size=2, stages=4, wait_d_free=True):
    sim = Sim(seed)
    r_full = Barrier("r_full", 1)
    t_full = [Barrier(f"t_full{s}", 1) for s in range(stages)]
    t_empty = [Barrier(f"t_empty{s}", 1) for s in range(stages)]
    x_full = [Barrier(f"x_full{g}", 1) for g in (0, 1)]
    x_free = [Barrier(f"x_free{g}", group_size) for g in (0, 1)]
    pb_full = [Barrier(f"pb_full{g}", group_size) for g in (0, 1)]
    d_free = [Barrier(f"d_free{g}", 1) for g in (0, 1)]
    done_bar = Barrier("done", 1)
    st = dict(T=[None] * stages, X=[None, None], X_loaded=[set(), set()], D=[None, None], D_written=[set(), set()], b_exec=-1,
              R=False)

    def tma():
        yield ("do", lambda: st.__setitem__("R", True))
        yield ("arrive", r_full)
        for i in range(n_t):
            s = i % stages
            yield ("wait", t_empty[s], ((i // stages) & 1) ^ 1, i // stages - 1)
            yield ("do", lambda i=i, s=s: st["T"].__setitem__(s, i))
            yield ("arrive", t_full[s])

    def mma():
        def exec_A(i):
            g = i & 1
            if not st["R"] or st["T"][i % stages] != i:
                raise ProtocolError(f"A({i}) reads stage holding {st['T'][i % stages]}")
            if st["X"][g] is not None and len(st["X_loaded"][g]) != group_size:
                raise ProtocolError(f"A({i}) overwrites X[{g}] (tile {st['X'][g]}) before the group loaded it")
            st["X"][g], st["X_loaded"][g] = i, set()

        def exec_B(i):
            g = i & 1
            if st["T"][i % stages] != i:
                raise ProtocolError(f"B({i}) reads stage holding {st['T'][i % stages]}")
            if st["D"][g] != i or len(st["D_written"][g]) != group_size:
                raise ProtocolError(f"B({i}) reads D[{g}] = tile {st['D'][g]} with {len(st['D_written'][g])}/{group_size} rows")
            st["b_exec"] = i

        def issue_A(i):
            s = i % stages
            yield ("wait", t_full[s], (i // stages) & 1, i // stages)
            yield ("push", ("mma", lambda i=i: exec_A(i)))
            yield ("push", ("commit", x_full[i & 1]))

        yield ("wait", r_full, 0, 0)
        yield from issue_A(0)
        if n_t > 1:
            yield from issue_A(1)
        for i in range(n_t):
            g = i & 1
            if i + 2 < n_t:
                yield ("wait", x_free[g], (i >> 1) & 1, i >> 1)
                yield from issue_A(i + 2)
            yield ("wait", pb_full[g], (i >> 1) & 1, i >> 1)
            yield ("push", ("mma", lambda i=i: exec_B(i)))
            yield ("push", ("commit", t_empty[i % stages]))
            yield ("push", ("commit", d_free[g]))
        yield ("push", ("commit", done_bar))

    def softmax(g, w):
        for i in range(g, n_t, 2):
            yield ("wait", x_full[g], (i >> 1) & 1, i >> 1)

            def load(i=i):
                if st["X"][g] != i:
                    raise ProtocolError(f"group {g} loads X holding tile {st['X'][g]} instead of {i}")
                st["X_loaded"][g].add(w)
            yield ("do", load)
            yield ("arrive", x_free[g])
            if wait_d_free and i >= 2:
                yield ("wait", d_free[g], ((i >> 1) - 1) & 1, (i >> 1) - 1)

            def write(i=i):
                if st["b_exec"] < i - 2:
                    raise ProtocolError(f"group {g} overwrites dS({i - 2}) before B({i - 2}) read it")
                if st["D"][g] != i:
                    st["D"][g], st["D_written"][g] = i, set()
                st["D_written"][g].add(w)
            yield ("do", write)
            yield ("arrive", pb_full[g])
        yield ("wait", done_bar, 0, 0)

    sim.add("tma", tma())
    sim.add("mma", mma())
    for g in (0, 1):
        for w in range(group_size):
            sim.add(f"soft{g}_{w}", softmax(g, w))
    sim.run()
    return sim.steps


# ---------------------------------------------------------------------------------------------------------------------
# persistent GEMM (gemm_tcgen05.cu): TMA ring of STAGES k-blocks, two TMEM accumulators, epilogue warps
# ---------------------------------------------------------------------------------------------------------------------
def run_gemm(seed, n_tiles=5, kb_per_tile=4, stages=3, n_epi=3, tiles_with_zero_k=()):
    """One worker (CTA or CTA pair): producer fills stage s when `empty[s]` says the MMAs that read it have retired;
    the MMA warp accumulates tile t into accumulator t & 1 once the epilogue has drained it (`tempty`), commits
    `empty[s]` per k-block and `tfull[acc]` per tile; every epilogue agent reads the accumulator and arrives `tempty`.
    `tiles_with_zero_k`: split-K slices without k-blocks (the kernel still runs the tfull / tempty handshake)."""
    sim = Sim(seed)
    full = [Barrier(f"full{s}", 1) for s in range(stages)]
    empty = [Barrier(f"empty{s}", 1) for s in range(stages)]
    tfull = [Barrier(f"tfull{a}", 1) for a in (0, 1)]
    tempty = [Barrier(f"tempty{a}", n_epi) for a in (0, 1)]
    st = dict(stage=[None] * stages, acc_tile=[None, None], acc_kb=[0, 0], acc_read=[set(), set()], consumed=-1)
    kbs = [0 if t in tiles_with_zero_k else kb_per_tile for t in range(n_tiles)]
    starts = [sum(kbs[:t]) for t in range(n_tiles)]

    def tma():
        n = 0
        for t in range(n_tiles):
            for kb in range(kbs[t]):
                s = n % stages
                yield ("wait", empty[s], ((n // stages) & 1) ^ 1, n // stages - 1)

                def load(n=n, s=s):
                    if st["stage"][s] is not None and st["consumed"] < st["stage"][s]:
                        raise ProtocolError(f"TMA overwrites stage {s} (k-block {st['stage'][s]}) before the MMA read it")
                    st["stage"][s] = n
                yield ("do", load)
                yield ("arrive", full[s])
                n += 1

    def mma():
        n = 0
        for t in range(n_tiles):
            a = t & 1
            yield ("wait", tempty[a], ((t >> 1) & 1) ^ 1, (t >> 1) - 1)
            for kb in range(kbs[t]):
                s = n % stages
                yield ("wait", full[s], (n // stages) & 1, n // stages)

                def exec_mma(n=n, s=s, t=t, a=a, kb=kb):
                    if st["stage"][s] != n:
                        raise ProtocolError(f"MMA reads stage {s} holding k-block {st['stage'][s]} instead of {n}")
                    if kb == 0:
                        if st["acc_tile"][a] is not None and len(st["acc_read"][a]) != n_epi:
                            raise ProtocolError(f"tile {t} overwrites accumulator {a} before the epilogue drained tile {st['acc_tile'][a]}")
                        st["acc_tile"][a], st["acc_kb"][a], st["acc_read"][a] = t, 0, set()
                    st["acc_kb"][a] += 1
                    st["consumed"] = n
                yield ("push", ("mma", exec_mma))
                yield ("push", ("commit", empty[s]))
                n += 1
            if kbs[t] == 0:
                def mark(t=t, a=a):
                    st["acc_tile"][a], st["acc_kb"][a], st["acc_read"][a] = t, 0, set()
                yield ("push", ("mma", mark))
            yield ("push", ("commit", tfull[a]))

    def epilogue(w):
        for t in range(n_tiles):
            a = t & 1
            yield ("wait", tfull[a], (t >> 1) & 1, t >> 1)

            def drain(t=t, a=a):
                if st["acc_tile"][a] != t or st["acc_kb"][a] != kbs[t]:
                    raise ProtocolError(f"epilogue {w} reads accumulator {a}: tile {st['acc_tile'][a]}, {st['acc_kb'][a]}/{kbs[t]} k-blocks")
                st["acc_read"][a].add(w)
            yield ("do", drain)
            yield ("arrive", tempty[a])

    sim.add("tma", tma())
    sim.add("mma", mma())
    for w in range(n_epi):
        sim.add(f"epi{w}", epilogue(w))
    sim.run()
    return sim.steps
