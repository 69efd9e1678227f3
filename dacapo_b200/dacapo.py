"""Python restatement of the DaCapo bootstrap planner (USENIX Security '24) on the Earth-level ops of
`dacapo_b200.earth`:

    RemoveBootstrap -> ScaleManagementUnit ids -> CandidateAnalysis (liveness, candidate cut points) -> BypassDetection
    -> CandidateSelection -> DaCapoPlanner (dynamic programming over cut points, every (from, to) segment scale-managed with
    PARS and costed with the target's latency table) -> BootstrapPlacement -> ProactiveRescaling -> EarlyModswitch

restated from lib/Dialect/Earth/Transforms/{RemoveBootstrap,BypassDetection,CandidateSelection,DaCapoPlanner,
CoverageRecorder,CodeSegmentation,BootstrapPlacement,LatencyEstimator}.cpp, lib/Dialect/Earth/Analysis/
{CandidateAnalysis,ScaleManagementUnit}.cpp and the `dacapo` pipeline of tools/optimizer.cpp:380-436.  The planner's
objective is the estimated latency on the TARGET, i.e. the per-op / per-level table of profiled_B200_GPU.json.

Differences from the MLIR implementation, by construction rather than by intent: a segment is cut out of the op list
directly instead of cloning the whole function and dead-code-eliminating it (CodeSegmentation.cpp), and BypassDetection
runs sequentially (the reference uses one std::thread per candidate edge, BypassDetection.cpp:45-60).
"""
from typing import Dict, List, Tuple

from . import earth
from .earth import Func, Params, PassFailed, V


# ------------------------------------------------------------------------------------------------ ScaleManagementUnit
def smu_ids(f: Func) -> Dict[int, int]:
    """SMUBuilder (ScaleManagementUnit.cpp:36-190): partition refinement of the values by the SMUs of their operands
    (forward) and of their users (backward) until the number of units is stable; non-consuming units are looked through.
    Only MulOp consumes scale (isConsume, EarthOps.td)."""
    vals = list(f.ops)
    consume = {id(v): v.kind == "mul" for v in vals}
    us = earth.users(f)
    ret_ids = {id(r) for r in f.rets}
    smu = {id(v): 0 for v in vals}
    st = {"idmax": 1}
    maps = {(True, False): {}, (False, False): {}, (True, True): {}, (False, True): {}}  # (forward, consume) -> {(nid, def): id}
    origin, definition, cdef = {}, {}, {}

    def key_of(v, forward):
        if forward:
            return frozenset(smu[id(o)] for o in v.ins) if v.kind != "arg" else frozenset()
        d = {smu[id(u)] for u in us[id(v)]}
        if id(v) in ret_ids:
            d.add(-1)  # the func.return user has no result: getID(op) == -1
        return frozenset(d)

    def define(v, forward):
        nid = smu[id(v)]
        if nid not in definition:
            d = key_of(v, forward)
            cdef[nid] = consume[id(v)]
            definition[nid] = d
            maps[(forward, consume[id(v)])][(nid, d)] = nid

    def look(v, forward):
        nid = smu[id(v)]
        d = key_of(v, forward)
        for defi in sorted(d):
            if (defi == nid or origin.get(defi) == nid) and not cdef.get(defi, False):
                sub = d - {defi}
                dd = definition.get(defi, frozenset())
                if sub <= dd:
                    d = dd
                    break
        m = maps[(forward, consume[id(v)])]
        k = (nid, d)
        if k not in m:
            m[k] = st["idmax"]
            definition[st["idmax"]] = d
            cdef[st["idmax"]] = consume[id(v)]
            origin[st["idmax"]] = nid
            st["idmax"] += 1
        smu[id(v)] = m[k]

    count = 0
    while count != len(set(smu.values())):
        count = len(set(smu.values()))
        for forward, order in ((True, vals), (False, vals[::-1])):
            origin.clear()
            definition.clear()
            for v in order:
                define(v, forward)
            for v in order:
                look(v, forward)
    return smu


# ------------------------------------------------------------------------------------------------ CandidateAnalysis
class ValueInfo:
    def __init__(self, opid, v):
        self.opid, self.v = opid, v
        self.live_outs: List[int] = []
        self.live_ins: List[int] = []
        self.valid_live_outs: List[int] = []
        self.dead_opid = -1
        self.threshold_opid = 1 << 62
        self.coverage = self.boot_coverage = -1


class CandidateAnalysis:
    """CandidateAnalysis.cpp:9-73: opids in program order, the ciphertexts live after every op, and an EDGE (candidate
    cut point) wherever a value is used by a different scale-management unit (ops 1..10 are excluded like the reference)."""

    def __init__(self, f: Func):
        self.f = f
        smu = smu_ids(f)
        us = earth.users(f)
        ret_ids = {id(r) for r in f.rets}
        last_use = {}
        for idx, o in enumerate(f.ops):
            for i in o.ins:
                last_use[id(i)] = idx
        for r in f.rets:
            last_use[id(r)] = len(f.ops)
        self.values: List[ValueInfo] = [ValueInfo(0, None)]
        self.edges = [0]
        self.users = us
        live_in: List[int] = []
        live_out: List[int] = []
        for idx, o in enumerate(f.ops):
            if o.kind == "arg":
                o.opid = -1
                continue
            opid = len(self.values)
            o.opid = opid
            self.values.append(ValueInfo(opid, o))
            if not o.cipher:
                continue
            for x in o.ins:
                if not x.cipher:
                    continue
                if last_use.get(id(x), -1) <= idx and x.opid in live_out:
                    live_out.remove(x.opid)
                    self.values[x.opid].dead_opid = opid
            live_out.append(opid)
            cross = any(smu[id(u)] != smu[id(o)] for u in us[id(o)]) or id(o) in ret_ids
            if cross and opid > 10:
                self.values[opid].live_outs = list(live_out)
                self.values[opid].live_ins = list(live_in)
                self.edges.append(opid)
            live_in = list(live_out)
        self.ret_opid = len(self.values)
        self.values.append(ValueInfo(self.ret_opid, None))
        self.to_from: Dict[int, List[int]] = {0: []}
        self.candidate_set: Dict[int, List[int]] = {}
        self.candidates: List[int] = []

    def use_opids(self, opid):
        return [u.opid for u in self.users[id(self.values[opid].v)] if u.opid >= 0]

    def is_bypass_edge(self, bp: int, to: int) -> bool:
        """ValueInfo::isBypassEdge (CandidateAnalysis.cpp:271-292): the live-out `bp` need not be bootstrapped at the cut
        `to` when every use it still has comes after the point where a bootstrap taken at `bp` would be spent anyway."""
        vi = self.values[bp]
        if vi.threshold_opid <= to:
            return True
        if bp == to:
            return False
        for u in self.use_opids(bp):
            if u <= vi.threshold_opid and to < u:
                return False
        return True

    def targets(self, opid: int, set_num: int) -> List[int]:
        if opid == self.ret_opid:
            return []
        vi = self.values[opid]
        return vi.live_outs if len(vi.live_outs) == set_num else vi.valid_live_outs

    def bypass_types(self, opid: int) -> List[bool]:
        return [self.is_bypass_edge(t, opid) for t in self.values[opid].live_outs]

    def sort_targets(self, set_num: int, opids=None) -> List[int]:
        out = []
        for a in (self.candidate_set.get(set_num, []) if opids is None else opids):
            for b in self.targets(a, set_num):
                if b not in out:
                    out.append(b)
        return out

    def finalize(self, set_num: int):
        c = [0]
        for i in range(1, set_num + 1):
            c += self.candidate_set.get(i, [])
        c.append(self.ret_opid)
        self.candidates = sorted(set(c))

    def push_from_coverage(self, frm: int, coverage: int, boot_coverage: int):
        vi = self.values[frm]
        vi.coverage, vi.boot_coverage = coverage, boot_coverage
        c = self.ret_opid if coverage < 0 else coverage
        bc = self.ret_opid if boot_coverage < 0 else boot_coverage
        for to in self.candidates:
            if to < bc and frm < to:
                self.to_from.setdefault(to, []).append(frm)
            elif to == self.ret_opid and c == self.ret_opid:
                self.to_from.setdefault(to, []).append(frm)


# ------------------------------------------------------------------------------------------------ segments
def _segment(ca: CandidateAnalysis, frm: int, to: int, boot_targets: List[int], in_types: List[Tuple[int, int]], nargs: int,
             walk_limit: int = 0) -> Tuple[Func, List[V]]:
    """BootstrapPlacement + CodeSegmentation for the ops (frm, to]: the function arguments and the values live after
    `frm` become arguments (typed by the plan that ends at `frm`), `boot_targets` among them are bootstrapped on entry,
    the values live after `to` are returned.  `walk_limit` > 0: only the ops (frm, frm + walk_limit] in program order,
    no returns and no dead-code elimination (forward walks that stop early, see _walk_until)."""
    f = ca.f
    m: Dict[int, V] = {}
    ops: List[V] = []
    arg_types = []
    argc = 0
    for o in f.ops:
        if o.kind == "arg":
            a = V("arg", (), argc)
            m[id(o)] = a
            ops.append(a)
            arg_types.append(in_types[argc])
            argc += 1
    for k, opid in enumerate(ca.values[frm].live_outs):
        a = V("arg", (), argc)
        ops.append(a)
        arg_types.append(in_types[nargs + k])
        argc += 1
        src = ca.values[opid].v
        if opid in boot_targets:
            b = V("boot", [a], 0)
            b.opid = opid
            ops.append(b)
            m[id(src)] = b
        else:
            m[id(src)] = a
    hi = to if not walk_limit else min(to, frm + walk_limit)
    import bisect
    if not hasattr(ca, "_opids"):
        ca._body = [o for o in f.ops if o.kind != "arg"]  # opids are increasing in program order
        ca._opids = [o.opid for o in ca._body]
    for o in ca._body[bisect.bisect_right(ca._opids, frm):bisect.bisect_right(ca._opids, hi)]:
        if any(id(i) not in m for i in o.ins):
            # an operand defined before the cut that is not live after it can only be a plaintext constant's user chain
            if all((not i.cipher) or id(i) in m for i in o.ins):
                for i in o.ins:
                    if id(i) not in m:
                        c = V(i.kind, (), i.attr, cipher=False)
                        ops.append(c)
                        m[id(i)] = c
            else:
                raise PassFailed(f"segment ({frm},{to}]: operand of opid {o.opid} is not live at the cut")
        n = V(o.kind, [m[id(i)] for i in o.ins], o.attr, o.cipher)
        n.opid = o.opid
        m[id(o)] = n
        ops.append(n)
    if walk_limit:
        return Func(ops, [], f.consts), arg_types
    if to == ca.ret_opid:
        rets = [m[id(r)] for r in f.rets]
    else:
        rets = [m[id(ca.values[t].v)] for t in ca.values[to].live_outs]
    seg = Func(ops, rets, f.consts)
    earth.dce(seg)
    return seg, arg_types


def _walk_until(ca, frm, boot_targets, in_types, nargs, P, pred):
    """PARS from the cut `frm` to the end of the function, stopping at the first op for which `pred(op)` holds
    (BypassDetection::findBypassEdge, CoverageRecorder)."""
    hit = {}

    def stop(op):
        r = pred(op)
        if r:
            hit["opid"] = op.opid
        return r

    # the stopping op is normally a few hundred ops away: walk a window first, the whole rest only if nothing stopped
    for window in (2000, 0):
        if window and frm + window >= ca.ret_opid:
            continue
        if window:
            seg, arg_types = _segment(ca, frm, ca.ret_opid, boot_targets, in_types, nargs, walk_limit=window)
        else:
            seg, arg_types = _segment(ca, frm, ca.ret_opid, boot_targets, in_types, nargs)
        hit.clear()
        done = {"end": False}
        try:
            earth.pars(seg, P, arg_types=arg_types, stop=stop)
            done["end"] = True
        except PassFailed:
            done["end"] = "opid" not in hit  # a failure before any stop: nothing further can be walked
            if not window or "opid" in hit:
                break
            continue
        if "opid" in hit or not window:
            break
    return hit.get("opid", -1)


# ------------------------------------------------------------------------------------------------ the passes
def bypass_detection(ca: CandidateAnalysis, P: Params, nargs: int, threshold=0.5):
    """BypassDetection.cpp:28-130: per candidate edge, bootstrap its live-outs, scale-manage forward and record the
    first multiply whose accumulated scale passes `threshold` of the modulus; then sort the edges by how many live-outs
    really need a bootstrap."""
    Rf = P.rescaling_factor
    over = lambda op: op.kind == "mul" and op.scale + op.level * Rf > P.boot_upper * Rf * threshold
    for frm in ca.edges:
        if frm == 0:
            continue
        lo = ca.values[frm].live_outs
        types = [(P.waterline, 0)] * nargs + [(Rf, 0)] * len(lo)
        t = _walk_until(ca, frm, lo, types, nargs, P, over)
        if t >= 0:
            ca.values[frm].threshold_opid = t
    for a in ca.edges:
        vi = ca.values[a]
        vi.valid_live_outs = [bp for bp in vi.live_outs if not ca.is_bypass_edge(bp, a)]
        ca.candidate_set.setdefault(len(vi.valid_live_outs), []).append(a)
        if len(vi.live_outs) != len(vi.valid_live_outs):
            ca.candidate_set.setdefault(len(vi.live_outs), []).append(a)


def place_bootstraps(f: Func, ca: CandidateAnalysis, targets: List[int]) -> Func:
    """BootstrapPlacement.cpp:24-50 on a clone: a bootstrap after every target value, all other uses redirected to it."""
    g, m = f.clone()
    tset = set(targets)
    out, sub = [], {}
    for o in g.ops:
        o.ins = [sub.get(id(i), i) for i in o.ins]
        out.append(o)
        if o.opid in tset and o.cipher:
            b = V("boot", [o], 0)
            b.opid = o.opid
            out.append(b)
            sub[id(o)] = b
    g.ops = out
    g.rets = [sub.get(id(r), r) for r in g.rets]
    return g


def candidate_selection(f: Func, ca: CandidateAnalysis, P: Params) -> int:
    """CandidateSelection.cpp:27-60: the smallest live-out count whose cut points, all bootstrapped, let PARS succeed."""
    mx = max(ca.candidate_set) if ca.candidate_set else 0
    for i in range(1, mx + 1):
        g = place_bootstraps(f, ca, ca.sort_targets(i))
        try:
            earth.pars(g, P)
        except PassFailed:
            continue
        ca.finalize(i)
        return i
    raise PassFailed("no candidate set makes the program fit the modulus chain")


def plan(f: Func, ca: CandidateAnalysis, P: Params, set_num: int, nargs: int, log=None):
    """DaCapoPlannerPass (DaCapoPlanner.cpp:39-219): best[to] = min over from of best[from] + latency(segment(from, to))."""
    Rf = P.rescaling_factor
    best = {0: (0.0, [], [(P.waterline, 0)] * nargs)}
    for to in ca.candidates:
        opt = None
        for frm in ca.to_from.get(to, []):
            if frm not in best:
                continue
            cost0, plan0, types0 = best[frm]
            tg = ca.targets(frm, set_num)
            try:
                seg, arg_types = _segment(ca, frm, to, tg, types0, nargs)
                mid = to != ca.ret_opid
                earth.pars(seg, P, arg_types=arg_types, mid_segment=mid, ret_bypass=ca.bypass_types(to) if mid else None)
                earth.early_modswitch(seg, P)
                earth.canonicalize(seg, P)
                cost = cost0 + earth.latency(seg, P)
            except PassFailed:
                continue
            if opt is None or cost < opt[0]:
                opt = (cost, plan0 + [to], [(P.waterline, 0)] * nargs + [(r.scale, r.level) for r in seg.rets])
        if opt is not None:
            best[to] = opt
        if to != ca.ret_opid and to in best:  # CoverageRecorder: how far a plan that cuts at `to` reaches
            types = best[to][2]
            tg = ca.targets(to, set_num)
            state = {"boot": -1}

            def pred(op, state=state):
                if op.kind != "mul":
                    return False
                acc = op.scale + op.level * Rf
                if state["boot"] < 0 and not acc < (P.boot_upper - P.boot_lower + 1) * Rf:
                    state["boot"] = op.opid
                    return False
                return not acc < P.boot_upper * Rf

            cov = _walk_until(ca, to, tg, types, nargs, P, pred)
            ca.push_from_coverage(to, cov, state["boot"])
        if log and to in best:
            log(f"  cut {to}: best latency {best[to][0] / 1e6:.4f} s with {len(best[to][1]) - 1} cuts")
    if ca.ret_opid not in best:
        raise PassFailed("the planner found no feasible plan")
    cost, cuts, _ = best[ca.ret_opid]
    return cost, ca.sort_targets(set_num, cuts), cuts


def compile_dacapo(graph, P: Params = None, log=None):
    """`hopts dacapo <waterline>` (optimizer.cpp:380-436).  Returns (program, func, report)."""
    P = P or Params()
    f = earth.from_graph(graph, keep_bootstraps=False)  # RemoveBootstrap
    nargs = sum(1 for o in f.ops if o.kind == "arg")
    ca = CandidateAnalysis(f)
    bypass_detection(ca, P, nargs)
    set_num = candidate_selection(f, ca, P)
    est, targets, cuts = plan(f, ca, P, set_num, nargs, log)
    g = place_bootstraps(f, ca, targets)
    earth.run_pars_pipeline(g, P)
    report = {"estimated_latency_s": est / 1e6, "bootstraps": len(targets), "cut_points": len(cuts) - 1, "selected_set": set_num,
              "candidate_edges": len(ca.edges) - 1, "candidates": len(ca.candidates), "estimated_latency_final_s": earth.latency(g, P) / 1e6}
    return earth.emit_hevm(g, P), g, report
