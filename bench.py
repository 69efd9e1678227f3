#!/usr/bin/env python
"""bench.py -- HEVM op micro-benchmark on B200 (BASELINE.json configs[1]).

Workload ("step"): one run() of a compiled-style HEVM program over a batch of independent
ciphertexts at the reference geometry (N = 2^15, 14 x 60-bit primes, SEAL_HEVM.cpp:39-53).
Every ciphertext walks the FULL modulus chain: for level l = 13 .. 1:
rotate (one key-switch, a different Galois key per ciphertext), multiply + relinearize,
rescale.  That is 13 rotates + 13 HMult+relin + 12 rescales (+1 register copy) per ciphertext.

  value : HEVM ops/s, run() only, inputs already encrypted and resident in HBM (what hc-test times)
  e2e   : same metric through the reference-facing C ABI with HOST buffers: encrypt() of the
          inputs (host doubles -> H2D -> encode -> public-key encrypt), run(), decrypt_result()
          (decode -> D2H) all inside the timed region
  --impl reference : the CPU restatement of the reference path (oracle/), one chain per host
          core, timed on the box's host cores (SEAL itself is not installable here)

One process per GPU; ranks are independent replicas over different ciphertexts (weak scaling,
no data-path collective: SURVEY.md 8e "independent ciphertexts").
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))
from dacapo_b200 import _binding, hevm_asm as asm  # noqa: E402


def _fixtures():
    """tests/fixtures.py: the committed ResNet-20 fixture and the path of the CPU oracle (checker / cpu_baseline leg only)."""
    sys.path.insert(0, str(REPO / "tests"))
    import fixtures
    return fixtures

LOGN, NPRIMES, TOP = 15, 14, 13
N = 1 << LOGN
SLOTS = N // 2
BYTES_LIMB = 8 * N
SEED = 0xDACA90
f64p = C.POINTER(C.c_double)


# ------------------------------------------------------------------------------------------------
def build_program(batch, top=TOP, check_after=3):
    """args 0..batch-1 stay pristine (encrypt() writes them); chains run in registers batch..2*batch-1."""
    p = asm.Program(init_level=top)
    args = [p.arg(60, top) for _ in range(batch)]
    work = [p.new_ct() for _ in range(batch)]
    mid = p.new_ct()
    steps = []
    for b in range(batch):
        k = b % 14
        steps.append((1 << k) if (b // 14) % 2 == 0 else -(1 << k))
    for b in range(batch):
        p.rotate(work[b], args[b], 0)  # register copy (rotate by 0)
    nks = 0
    for i, l in enumerate(range(top, 0, -1)):
        for b in range(batch):
            p.rotate(work[b], work[b], steps[b])
            p.emit(asm.MULCC, work[b], work[b], work[b])
            nks += 2
            if l > 1:
                p.emit(asm.RESCALE, work[b], work[b])
        if i + 1 == check_after:
            p.rotate(mid, work[0], 0)
    for b in range(batch):
        p.result(work[b], 60, 1)
    p.result(mid, 60, top - check_after)
    return p, steps, nks


def expected_mid(x, step, rounds):
    v = x.copy()
    for _ in range(rounds):
        v = np.roll(v, -step)
        v = v * v
    return v


def make_vm(lib, keydir):
    if not os.path.isfile(os.path.join(keydir, "hevm_params.bin")):
        os.environ.update(HEVM_LOGN=str(LOGN), HEVM_NUM_PRIMES=str(NPRIMES), HEVM_SEED=str(SEED), HEVM_PRIME_BITS="60")
        lib.create_context(keydir.encode())
    return lib.initFullVM(keydir.encode(), True)


def load_program(lib, vm, prog, tmp):
    cst, hv = os.path.join(tmp, "bench.cst"), os.path.join(tmp, "bench.hevm")
    prog.save(cst, hv)
    lib.load(vm, cst.encode(), hv.encode())
    lib.preprocess(vm)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path).read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def measured_peak():
    try:
        return float(json.loads((REPO / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def cpu_chain_rate(nthreads, reps, keydirs):
    """Oracle (CPU port of the reference path): `nthreads` independent VMs, each runs one full chain per rep."""
    olib = _binding.bind(_fixtures().ORACLE_LIB)
    prog, steps, _ = build_program(1)
    nops = len(prog.ops)
    tmp = tempfile.mkdtemp(prefix="hevm_cpu_")
    vms = [None] * nthreads

    def init(i):
        vms[i] = make_vm(olib, keydirs[i])
        load_program(olib, vms[i], prog, tempfile.mkdtemp(prefix=f"hevm_cpu{i}_", dir=tmp))
        x = np.random.default_rng(i).uniform(-0.5, 0.5, SLOTS)
        olib.encrypt(vms[i], 0, x.ctypes.data_as(f64p), SLOTS)

    ts = [threading.Thread(target=init, args=(i,)) for i in range(nthreads)]
    [t.start() for t in ts]
    [t.join() for t in ts]

    def one_step():
        th = [threading.Thread(target=olib.run, args=(vms[i],)) for i in range(nthreads)]
        t0 = time.perf_counter()
        [t.start() for t in th]
        [t.join() for t in th]
        return time.perf_counter() - t0

    return one_step, nops * nthreads


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = max(1, min(os.cpu_count() or 1, 16))  # one oracle VM (2.9 GB of keys) per core
    subprocess.run(["make", "-s", "-C", str(REPO / "oracle")], check=True)
    base = tempfile.mkdtemp(prefix="hevm_ref_keys_")
    keydirs = []
    for i in range(cores):
        d = os.path.join(base, str(i))
        os.makedirs(d)
        keydirs.append(d)
    one_step, ops_per_step = cpu_chain_rate(cores, 1, keydirs)
    for _ in range(args.warmup):
        one_step()
    t = [one_step() for _ in range(args.steps)]
    total = sum(t)
    val = ops_per_step * args.steps / total
    line = {
        "impl": "reference", "metric": "hevm_ops_per_s", "value": val, "unit": "ops/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args.batch),
        "cpu_baseline": {"value": val, "unit": "ops/s", "cores": cores, "kind": "port",
                         "sample": f"bounded sample of the workload: {cores} of its independent ciphertext chains per step, one per host core "
                                   f"({ops_per_step // cores} HEVM ops each); CPU restatement of the SEAL 4.0 evaluator (oracle/), SEAL itself is not installable offline"},
        "e2e": {"value": val, "unit": "ops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def workload_config(batch, note=None):
    c = {"workload": "HEVM op microbench: per ciphertext, levels 13..1: rotate(1 key-switch) + HMult+relin + rescale "
                     "(N=2^15, 14x60-bit primes, full modulus chain)",
         "batch_ciphertexts_per_gpu": batch, "poly_degree": N, "num_primes": NPRIMES,
         "l2": "working set per step (>= 1.3 GB of distinct key-switch keys) exceeds the 126 MB L2; no explicit flush"}
    if note:
        c["note"] = note
    return c


def resnet20_real(lib, vm, tmp, reps=3, variant="", cpu=False):
    """The real encrypted ResNet-20 (BASELINE.json configs[0]) from the committed fixture: program traced from the
    reference's examples/benchmarks/ResNet.py, compiled by dacapo_b200.compiler; rms against the plaintext torch
    logits exactly like examples/tests/ResNet.py:113-118 (hc-test times run() only)."""
    fixtures = _fixtures()
    cst, hv, x, expected, meta = fixtures.resnet20_files(tmp, variant)
    if variant:  # another ring size (nt = 2^16 slots => N = 2^17): its own VM, keys derived on the device
        kd = tempfile.mkdtemp(prefix="hevm_bench_keys_big_")
        os.environ.update(HEVM_LOGN=str(meta["logN"]), HEVM_NUM_PRIMES=str(NPRIMES), HEVM_SEED=str(SEED), HEVM_PRIME_BITS="60")
        lib.create_context(kd.encode())
        vm = lib.initFullVM(kd.encode(), True)
        os.environ.update(HEVM_LOGN=str(LOGN))
    t0 = time.perf_counter()
    lib.load(vm, cst.encode(), hv.encode())
    lib.preprocess(vm)
    t_pre = time.perf_counter() - t0
    out = np.zeros(meta.get("slots", SLOTS))
    lat, e2e, first_run, capture_run = [], [], None, None
    for i in range(reps + 2):
        t0 = time.perf_counter()
        lib.encrypt(vm, 0, x.ctypes.data_as(f64p), x.size)
        t1 = time.perf_counter()
        lib.run(vm)
        t2 = time.perf_counter()
        lib.decrypt_result(vm, 0, out.ctypes.data_as(f64p))
        t3 = time.perf_counter()
        if i == 0:
            first_run = t2 - t1    # the cold run() of a process: the schedule is issued on the lanes, no graph yet
        elif i == 1:
            capture_run = t2 - t1  # second run(): stream capture + graph instantiate + launch
        else:
            lat.append(t2 - t1)
            e2e.append(t3 - t0)
    res = out[:meta["n_out"]] * meta["post_scale"]
    err = res - expected
    extra = {}
    if cpu and not variant:
        est, _ = cpu_port_estimate(program_histogram(hv))
        extra = {"cpu_port_estimate_s": est, "speedup_vs_cpu_port_estimate": est / float(np.median(lat)),
                 "cpu_port_estimate_note": "this very program on the CPU port (oracle, one thread): sum over (op, level) of count x measured "
                                           "oracle time for key-switch steps, mulcc, rescale, mulcp, addcc and bootstrap"}
    return {**extra, "what": f"encrypted ResNet-20 (SiLU, nt=2^{int(np.log2(meta.get('slots', SLOTS)))} slots, N=2^{meta.get('logN', LOGN)}, 14x60-bit primes, waterline 40), "
                    "synthetic seeded input, weights examples/data/resnet20.silu.model; program compiled by dacapo_b200.compiler "
                    "(not hecate-opt), bootstrap levels from the measured cost profile",
            "run_latency_s": float(np.median(lat)), "first_run_s": first_run, "graph_capture_run_s": capture_run,
            "latency_note": "run_latency_s = warm run() (CUDA-graph replay, third run on); first_run_s = the first run() of the process "
                            "(issued on the lanes without a graph), which is what one `hc-test` invocation of the reference times "
                            "(examples/tests/ResNet.py:109-111: a single cold run() per process); graph_capture_run_s = the second run(), "
                            "which captures and instantiates the graph",
            "e2e_latency_s": float(np.median(e2e)), "load_preprocess_s": t_pre,
            "rms": float(np.sqrt(np.sum(err * err) / res.shape[-1])), "argmax_ok": bool(np.argmax(res) == np.argmax(expected)),
            "lowered_ops": meta["lowered_ops"], "hevm_ops": meta["hevm_ops"],
            "reference_README": {"latency_s": 53.726, "rms": 9.515e-4, "hardware": "unspecified CPU, SEAL single thread (README.md:186-187)"},
            "speedup_vs_README_latency": 53.726 / float(np.median(lat)), "speedup_vs_README_latency_first_run": 53.726 / first_run}


_CPU_OP_CACHE = {}


def cpu_port_estimate(per_level):
    """Latency of an op mix on the CPU port (oracle, one thread): every (op, level) pair is timed once on the oracle and
    multiplied by its count.  per_level: {(name, level): count}, names ks_steps / mulcc / rescale / mulcp / addcc / bootstrap
    (for bootstrap the level is the target level)."""
    olib = _binding.bind(_fixtures().ORACLE_LIB)
    if "vm" not in _CPU_OP_CACHE:
        kd = tempfile.mkdtemp(prefix="hevm_cpu_mix_")
        ovm = make_vm(olib, kd)
        olib.hevmx_resize(ovm, 4, 2)
        _CPU_OP_CACHE["vm"] = ovm
    ovm = _CPU_OP_CACHE["vm"]
    u64p = C.POINTER(C.c_uint64)
    primes = np.zeros(NPRIMES, dtype=np.uint64)
    olib.hevmx_primes(ovm, primes.ctypes.data_as(u64p))
    est, table = 0.0, {}
    rng = np.random.default_rng(3)
    ops = (("ks_steps", asm.ROTATE, 1), ("mulcc", asm.MULCC, 0), ("rescale", asm.RESCALE, 0), ("mulcp", asm.MULCP, 0),
           ("addcc", asm.ADDCC, 0), ("bootstrap", asm.BOOTSTRAP, None))
    for lvl in sorted({l for (_, l) in per_level}):
        for name, opc, rhs in ops:
            n = per_level.get((name, lvl), 0)
            if not n:
                continue
            key = (name, lvl)
            if key not in _CPU_OP_CACHE:
                src_l = 2 if name == "bootstrap" else lvl
                a = np.zeros((2, src_l, N), dtype=np.uint64)
                for i in range(src_l):
                    a[:, i, :] = rng.integers(0, int(primes[i]), size=(2, N), dtype=np.uint64)
                olib.hevmx_ct_write(ovm, 0, a.ctypes.data_as(u64p), src_l, 2.0 ** 40)
                olib.hevmx_pt_write(ovm, 0, a[0].ctypes.data_as(u64p), src_l, 2.0 ** 40)
                t0 = time.perf_counter()
                olib.hevmx_exec(ovm, opc, 1, 0, lvl if rhs is None else rhs)
                _CPU_OP_CACHE[key] = time.perf_counter() - t0
            table[f"{name}@{lvl}"] = _CPU_OP_CACHE[key]
            est += n * _CPU_OP_CACHE[key]
    return est, table


def program_histogram(hevm_path):
    """(op, level) counts of a compiled program (rotations are single key-switch steps in our compiler's output)."""
    p = asm.parse_hevm(open(hevm_path, "rb").read())
    names = {asm.ROTATE: "ks_steps", asm.MULCC: "mulcc", asm.RESCALE: "rescale", asm.MULCP: "mulcp", asm.ADDCC: "addcc", asm.BOOTSTRAP: "bootstrap"}
    lv, hist = {i: lvl for i, lvl in enumerate(p.arg_level)}, {}
    for oc, d, l, r in p.ops:
        if oc in (asm.ENCODE, asm.PLACEHOLDER):
            continue
        L = lv.get(l, p.init_level)
        res = L - 1 if oc == asm.RESCALE else L - r if oc == asm.MODSWITCH else r if oc == asm.BOOTSTRAP else L
        if oc in names:
            key = (names[oc], res if oc == asm.BOOTSTRAP else L)
            hist[key] = hist.get(key, 0) + 1
        lv[d] = res
    return hist


def resnet_mix(lib, vm, tmp, cpu=True, reps=3):
    """ResNet-20 op-mix replay (SURVEY 8d fallback; dacapo_b200/workloads.py): run() latency on the GPU and the
    CPU port's time for the same op mix, estimated from its per-op / per-level timings."""
    from dacapo_b200 import workloads
    prog, stats, per_level = workloads.resnet20_opmix()
    cst, hv = os.path.join(tmp, "resnet_mix.cst"), os.path.join(tmp, "resnet_mix.hevm")
    prog.save(cst, hv)
    t0 = time.perf_counter()
    lib.load(vm, cst.encode(), hv.encode())
    lib.preprocess(vm)
    t_pre = time.perf_counter() - t0
    x = np.random.default_rng(5).uniform(-0.5, 0.5, SLOTS)
    out = np.zeros(SLOTS)
    lat, e2e = [], []
    for i in range(reps + 1):
        t0 = time.perf_counter()
        lib.encrypt(vm, 0, x.ctypes.data_as(f64p), SLOTS)
        t1 = time.perf_counter()
        lib.run(vm)
        t2 = time.perf_counter()
        lib.decrypt_result(vm, 0, out.ctypes.data_as(f64p))
        t3 = time.perf_counter()
        if i:  # (run 0 is issued on the lanes, run 1 captures the CUDA graph: the median below is a replay)
            lat.append(t2 - t1)
            e2e.append(t3 - t0)
    res = {"what": "ResNet-20 op-mix replay, NOT the compiled network (hecate-opt/MLIR unavailable): op counts of "
                   "examples/benchmarks/ResNet.py at nt=2^14 (SURVEY App. C), 19 bootstrap segments, levels 13..3",
           "ops": {k: v for k, v in stats.items()}, "hevm_ops": len(prog.ops), "run_latency_s": float(np.median(lat)),
           "e2e_latency_s": float(np.median(e2e)), "preprocess_s": t_pre, "finite_output": bool(np.all(np.isfinite(out))),
           "reference_README_latency_s": 53.726}
    if cpu:
        est, _ = cpu_port_estimate(per_level)
        res["cpu_port_estimate_s"] = est
        res["cpu_port_estimate_note"] = ("sum over (op, level) of count x single-thread oracle time for rotate-step, mulcc, rescale, "
                                         "mulcp, addcc, bootstrap (negate / addcp / modswitch not counted)")
        res["speedup_vs_cpu_port_estimate"] = est / res["run_latency_s"]
    return res


def measure_traffic_ncu(level=TOP, timeout=240):
    """LIVE DRAM traffic of one rotate at `level`: a short ncu pass over tools/ks_probe.py with --cache-control none (L2
    residency as in a real run: the probe's warm-up ops ran just before, with other keys) -- dram__bytes_read + write
    per kernel of the profiled rotate.  Returns None when ncu is not usable on this box."""
    import csv
    import io
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.isfile(ncu):
        return None
    log = tempfile.mktemp(prefix="ncu_traffic_", suffix=".csv")
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--cache-control", "none",
           "--clock-control", "none", "--profile-from-start", "off", "--print-units", "base", "--csv", "--log-file", log,
           sys.executable, str(REPO / "tools" / "ks_probe.py"), str(level)]
    env = dict(os.environ, HEVM_FUSED="0")
    try:
        subprocess.run(cmd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=timeout, check=True)
        rows = list(csv.reader(io.StringIO(open(log).read())))
    except Exception as e:  # no profiling permission, ncu missing, timeout ...
        return {"error": f"{type(e).__name__}"}
    hdr = next((i for i, r in enumerate(rows) if r and r[0] == "ID"), None)
    if hdr is None:
        return {"error": "no ncu rows"}
    ix = {h: i for i, h in enumerate(rows[hdr])}
    kern = {}
    for r in rows[hdr + 1:]:
        if len(r) < len(ix):
            continue
        k = kern.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]].split("(")[0]})
        try:
            k[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            pass
    ids = sorted(kern)
    rot = [kern[i] for i in ids[:5]]  # the probe's first profiled op is the rotate: 5 launches
    if len(rot) < 5 or "k_mac" not in rot[2]["name"]:
        return {"error": "unexpected launch list", "names": [k["name"] for k in rot]}
    byt = lambda k: k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0)
    return {"level": level, "cache_control": "none", "k_mac_dram_bytes": byt(rot[2]), "rotate_dram_bytes": sum(byt(k) for k in rot),
            "per_kernel": [{"kernel": k["name"].replace("void ", ""), "dram_bytes": byt(k), "ns_under_ncu": k.get("gpu__time_duration.sum")} for k in rot]}


def batch_sweep(lib, vm, peak, levels=(1, 4, 8, 13), batches=(1, 8, 64, 256)):
    """BASELINE.json configs[1] "batched ciphertexts": n independent ciphertexts per launch (hevmx_exec_batch, one
    persistent kernel per op class over the whole batch; n = 1 is the ordinary op-by-op path) -- microseconds per op
    and the fraction of the HBM roofline on SURVEY 8(d) algorithmic bytes (key bytes counted per ciphertext)."""
    from dacapo_b200 import profile
    u64p, i64p = C.POINTER(C.c_uint64), C.POINTER(C.c_int64)
    nmax = max(batches)
    primes = np.zeros(NPRIMES, dtype=np.uint64)
    lib.hevmx_primes(vm, primes.ctypes.data_as(u64p))
    lib.hevmx_resize(vm, 2 * nmax, 1)
    rng = np.random.default_rng(11)
    out = {}
    for l in levels:
        a = np.zeros((2, l, N), dtype=np.uint64)
        for i in range(l):
            a[:, i, :] = rng.integers(0, int(primes[i]), size=(2, N), dtype=np.uint64)
        for r in range(nmax):
            lib.hevmx_ct_write(vm, r, a.ctypes.data_as(u64p), l, 2.0 ** 40)
        for name, opc in (("rotate", asm.ROTATE), ("mulcc", asm.MULCC), ("rescale", asm.RESCALE)):
            if opc == asm.RESCALE and l < 2:
                continue
            row = {}
            for n in batches:
                src = np.arange(n, dtype=np.int64)
                dst = src + nmax
                steps = (1, -2, 4, -8, 16, -32, 64, -128)  # eight different Galois keys per batch: no key sits in L2 by construction
                rhs = src.copy() if opc == asm.MULCC else (np.array([steps[k % 8] & 0xFFFF for k in range(n)], dtype=np.int64) if opc == asm.ROTATE else np.zeros(n, dtype=np.int64))
                reps = max(2, 128 // n)
                call = lambda: lib.hevmx_exec_batch(vm, opc, n, dst.ctypes.data_as(i64p), src.ctypes.data_as(i64p), rhs.ctypes.data_as(i64p))
                call()
                lib.hevmx_sync(vm)
                lib.hevmx_timer(vm, 0)
                for _ in range(reps):
                    call()
                us = lib.hevmx_timer(vm, 1) * 1e3 / (reps * n)
                row[str(n)] = {"us_per_op": round(us, 2), "frac": round(profile.algorithmic_bytes(name, l) / (us * 1e-6) / 1e9 / peak, 4)}
            out.setdefault(name, {})[str(l)] = row
    return out


def oracle_resnet_subprocess():
    """Starts the CPU port (oracle, one thread) on the committed ResNet-20 program in a separate process: its run() time is
    MEASURED (~70 s on the GPU box's host) while the GPU sections of the bench proceed; joined at the end."""
    code = (
        "import sys, time, json, ctypes as C, tempfile\n"
        f"sys.path.insert(0, {str(REPO)!r}); sys.path.insert(0, {str(REPO / 'tests')!r})\n"
        "import numpy as np, fixtures\n"
        "from dacapo_b200 import _binding\n"
        "from util import make_vm\n"
        "lib = _binding.bind(fixtures.ORACLE_LIB)\n"
        "cst, hv, x, expected, meta = fixtures.resnet20_files(tempfile.mkdtemp())\n"
        "vm, _ = make_vm(lib, 15, 14)\n"
        "t = time.perf_counter(); lib.load(vm, cst.encode(), hv.encode()); lib.preprocess(vm); tp = time.perf_counter() - t\n"
        "f64p = C.POINTER(C.c_double)\n"
        "lib.encrypt(vm, 0, x.ctypes.data_as(f64p), x.size)\n"
        "t = time.perf_counter(); lib.run(vm); tr = time.perf_counter() - t\n"
        "out = np.zeros(1 << 14); lib.decrypt_result(vm, 0, out.ctypes.data_as(f64p))\n"
        "res = out[:meta['n_out']] * meta['post_scale']\n"
        "print(json.dumps({'run_s': tr, 'load_preprocess_s': tp, 'rms': float(np.sqrt(np.sum((res - expected) ** 2) / res.shape[-1]))}))\n")
    subprocess.run(["make", "-s", "-C", str(REPO / "oracle")], check=True)
    return subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=28, help="independent ciphertexts per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--op-table", action="store_true", help="(default at N=1) measure the per-op/per-level table and emit profiled_B200_GPU.json")
    ap.add_argument("--no-op-table", action="store_true")
    ap.add_argument("--no-resnet-mix", action="store_true", help="skip the ResNet-20 op-mix replay (N=1 only)")
    ap.add_argument("--no-sharded", action="store_true", help="skip the limb-sharded key-switch section (N>1 only)")
    ap.add_argument("--ncu-range", action="store_true",
                    help="bracket the timed region with cudaProfilerStart/Stop (ncu --profile-from-start off captures exactly the timed launches)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libB200_HEVM.so has no CPU path")
    torch.cuda.set_device(local_rank)
    os.environ["HEVM_DEVICE"] = str(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        from dacapo_b200 import dist as D
        return D.max_over_ranks(x, device="cuda")

    lib = _binding.bind(_binding.B200_LIB)  # raises if the CUDA library is missing
    oracle_proc = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.no_resnet_mix:
        oracle_proc = oracle_resnet_subprocess()  # CPU port of the ResNet-20 program, measured while the GPU sections run
    keydir = tempfile.mkdtemp(prefix=f"hevm_bench_keys_r{rank}_")
    vm = make_vm(lib, keydir)
    prog, steps, nks = build_program(args.batch)
    nops = len(prog.ops)
    tmp = tempfile.mkdtemp(prefix="hevm_bench_")
    load_program(lib, vm, prog, tmp)

    # inputs / outputs in pinned host memory
    nres = args.batch + 1
    xin = torch.empty((args.batch, SLOTS), dtype=torch.float64).pin_memory()
    xout = torch.empty((nres, SLOTS), dtype=torch.float64).pin_memory()
    rng = np.random.default_rng(1 + rank)
    xin.numpy()[:] = rng.uniform(-0.5, 0.5, (args.batch, SLOTS))
    xin_np, xout_np = xin.numpy(), xout.numpy()

    def encrypt_all():
        for b in range(args.batch):
            lib.encrypt(vm, b, xin_np[b].ctypes.data_as(f64p), SLOTS)

    def decrypt_all():
        for r in range(nres):
            lib.decrypt_result(vm, r, xout_np[r].ctypes.data_as(f64p))

    # ---- correctness gate: the check register must match plain numpy ----
    encrypt_all()
    lib.run(vm)
    decrypt_all()
    exp = expected_mid(xin_np[0], steps[0], 3)
    err = float(np.max(np.abs(xout_np[args.batch] - exp)))
    if not err < 1e-4:
        raise SystemExit(f"bench: decrypted check value differs from plain maths (max err {err})")

    # ---- value: run() only, inputs resident in HBM; CUDA events on the VM stream ----
    for _ in range(args.warmup):
        lib.run(vm)
    launches0 = lib.hevmx_param(vm, 6)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    if args.ncu_range:
        lib.hevmx_profiler_range(vm, 1)
    lib.hevmx_timer(vm, 0)
    for _ in range(args.steps):
        lib.run(vm)
    ms = lib.hevmx_timer(vm, 1)
    if args.ncu_range:
        lib.hevmx_profiler_range(vm, 0)
    barrier()
    clocks = sampler.stop()
    launches = lib.hevmx_param(vm, 6) - launches0
    ms = max_over_ranks(ms)
    value = nops * world * args.steps / (ms / 1e3)

    # ---- e2e: encrypt (H2D) + run + decrypt (D2H) through the C ABI, wall clock, max over ranks ----
    for _ in range(2):
        encrypt_all(), lib.run(vm), decrypt_all()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        encrypt_all()
        lib.run(vm)
        decrypt_all()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = nops * world * args.steps / e2e_s

    # ---- roofline of the dominant kernel: per-class CUDA-event timing over one more step ----
    lib.hevmx_profile(vm, 1)
    lib.run(vm)
    lib.hevmx_profile(vm, 0)
    kern = {}
    cls = 0
    while True:
        t, c = C.c_double(), C.c_int64()
        name = lib.hevmx_profile_read(vm, cls, C.byref(t), C.byref(c))
        if name is None:
            break
        if c.value:
            kern[name.decode()] = {"ms": t.value, "launches": c.value}
        cls += 1
    total_kernel_ms = sum(k["ms"] for k in kern.values())
    dom = max(kern, key=lambda k: kern[k]["ms"])
    peak, peak_src = measured_peak()
    # algorithmic bytes of one fwd_B_mac launch at level l: digits l(l+1) + key 2l(l+1) + out 2(l+1) limbs
    alg = {"fwd_B_mac": lambda l: (3 * l * l + 5 * l + 2) * BYTES_LIMB,
           "fwd_A_modup": lambda l: (l + l * l) * BYTES_LIMB}
    roof = None
    if dom in alg:
        launches_per_level = 2 * args.batch  # rotate + mulcc per ciphertext per level
        tot_bytes = sum(alg[dom](l) for l in range(1, TOP + 1)) * launches_per_level
        ach = tot_bytes / (kern[dom]["ms"] / 1e3) / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": None, "peak_source": peak_src, "share_of_step": kern[dom]["ms"] / total_kernel_ms,
                "avg_launch_us": 1e3 * kern[dom]["ms"] / kern[dom]["launches"]}
    # op-level roofline per SURVEY 8(d): key-switch bytes (2l^2+6l)B (rotate) / (2l^2+8l)B (HMult+relin), rescale (4l-2)B
    op_bytes = args.batch * sum((2 * l * l + 6 * l) + (2 * l * l + 8 * l) + ((4 * l - 2) if l > 1 else 0) for l in range(1, TOP + 1)) * BYTES_LIMB
    op_roof = {"algorithmic_GB_per_step": op_bytes / 1e9, "achieved_GBps": op_bytes * args.steps / (ms / 1e3) / 1e9 / 1.0,
               "frac_of_peak": op_bytes * args.steps / (ms / 1e3) / 1e9 / peak}

    line = {
        "metric": "hevm_ops_per_s", "value": value, "unit": "ops/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": workload_config(args.batch),
        "e2e": {"value": e2e, "unit": "ops/s", "h2d_bytes_per_step": args.batch * SLOTS * 8, "d2h_bytes_per_step": nres * SLOTS * 8,
                "ms_per_step": 1e3 * e2e_s / args.steps},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "op_roofline": op_roof,
        "keyswitch_per_s": nks * world * args.steps / (ms / 1e3), "hevm_ops_per_step": nops, "check_max_err": err,
    }
    details = {"kernels": kern}  # bulky tables go to bench_details.json (the JSON line stays readable in a log tail)

    if rank == 0 and world == 1 and not args.no_resnet_mix:
        r20 = resnet20_real(lib, vm, tmp, cpu=not args.no_cpu_baseline)
        details["resnet20"] = dict(r20)
        line["resnet20"] = {k: r20[k] for k in ("run_latency_s", "first_run_s", "e2e_latency_s", "rms", "argmax_ok", "hevm_ops", "speedup_vs_README_latency",
                                                 "speedup_vs_README_latency_first_run", "cpu_port_estimate_s") if k in r20}
        r16 = resnet20_real(lib, vm, tempfile.mkdtemp(prefix="hevm_bench_nt16_"), reps=2, variant="_nt16")
        details["resnet20_nt16"] = r16
        line["resnet20_nt16"] = {k: r16[k] for k in ("run_latency_s", "first_run_s", "rms", "argmax_ok", "hevm_ops")}
        details["resnet20_opmix"] = resnet_mix(lib, vm, tmp, cpu=not args.no_cpu_baseline)

    if (args.op_table or world == 1) and not args.no_op_table and rank == 0:
        from dacapo_b200 import profile
        table, prof = profile.emit_profile(str(REPO / "profiled_B200_GPU.json"), lib, vm)
        kl13 = table.pop("_kernels_rotate_l13", None)
        details["op_table_us"] = {op: {str(l): round(v, 2) for l, v in lv.items()} for op, lv in table.items()}
        # the same ops with 32 independent instances in flight through run() (profile.measure_throughput_table): the
        # table the bootstrap planner of dacapo_b200.compiler minimises
        details["op_table_throughput_us"] = prof.get("latencyTableThroughput")
        if kl13 and "fwd_B_mac" in kl13:
            # roofline of the dominant kernel on ONE well-defined launch shape: the key-switch MAC kernel at level 13.
            # Algorithmic (compulsory) bytes of that launch, SURVEY 8(d) style: the key 2l(l+1)B, the NTT-form target it
            # multiplies on the diagonal lB and the accumulators it writes 2(l+1)B.  The half-transformed mod-up matrix
            # l(l+1)B that the previous kernel hands over is NOT compulsory traffic and is reported separately.
            l = TOP
            alg_bytes = (2 * l * (l + 1) + l + 2 * (l + 1)) * BYTES_LIMB
            inter_bytes = l * (l + 1) * BYTES_LIMB
            us = 1e3 * kl13["fwd_B_mac"]["ms"] / kl13["fwd_B_mac"]["launches"]
            tr = measure_traffic_ncu(l)  # LIVE: ncu pass with --cache-control none inside this run
            details["traffic_ncu"] = tr
            tot = sum(k["ms"] for k in kl13.values())
            details["roofline_all_levels"] = line["roofline"]
            line["roofline"] = {"bound": "hbm", "kernel": "k_mac (key-switch inner product + forward pass B), level 13", "achieved": alg_bytes / (us * 1e-6) / 1e9,
                                "peak": peak, "unit": "GB/s", "frac": alg_bytes / (us * 1e-6) / 1e9 / peak,
                                "traffic": (tr or {}).get("k_mac_dram_bytes"), "traffic_source": "ncu --cache-control none, measured in this run" if tr and "k_mac_dram_bytes" in tr else f"unavailable: {tr}",
                                "algorithmic_bytes_per_launch": alg_bytes, "intermediate_bytes_per_launch": inter_bytes,
                                "frac_with_intermediate": (alg_bytes + inter_bytes) / (us * 1e-6) / 1e9 / peak,
                                "avg_launch_us": us, "share_of_rotate_l13": kl13["fwd_B_mac"]["ms"] / tot,
                                "rotate_l13_dram_bytes": (tr or {}).get("rotate_dram_bytes"), "rotate_l13_algorithmic_bytes": profile.algorithmic_bytes("rotate", l),
                                "peak_source": peak_src,
                                "ceiling_note": "issue-bound, not HBM-bound: IMAD / LOP3 / SHF issue at 0.5 per clk per SMSP and do not overlap (tools/pipe_bench2.cu), "
                                                "a 64-bit Shoup butterfly costs 39 clk per warp => >= 0.236 us per limb-NTT vs 0.080 us at the HBM roofline"}
            details["kernels_rotate_l13"] = kl13
        line["op_roofline_l13"] = {op: {"us": round(table[op][13], 2), "frac": round(profile.algorithmic_bytes(op, 13) / (table[op][13] * 1e-6) / 1e9 / peak, 4)}
                                   for op in ("rotate", "mulcc", "rescale", "addcc", "mulcp")}
        sweep = batch_sweep(lib, vm, peak)
        details["batch_sweep"] = sweep
        line["batch_sweep_us_per_op"] = {op: {l: {n: v["us_per_op"] for n, v in row.items()} for l, row in lv.items()} for op, lv in sweep.items()}
        line["op_roofline_by_batch_l13"] = {op: {n: v["frac"] for n, v in lv["13"].items()} for op, lv in sweep.items()}

    if rank == 0 and world == 1 and not args.no_op_table:
        # stand-alone NTT / INTT throughput on device-resident limbs (metric row "NTT/INTT ops/s vs HBM roofline"):
        # 182 limbs (one launch pair) under each of the 13 data primes = 2 366 limbs = 620 MB per pass; algorithmic bytes
        # 2 x 8N per limb
        nb, npr = NPRIMES * (NPRIMES - 1), TOP
        ntt = {}
        for name, inv in (("ntt", 0), ("intt", 1)):
            ms_pass = lib.hevmx_ntt_bench(vm, nb, npr, inv, 5)
            rate = nb * npr / (ms_pass * 1e-3)
            ntt[name] = {"limb_transforms_per_s": rate, "us_per_limb": 1e6 / rate, "achieved_GBps": rate * 2 * BYTES_LIMB / 1e9,
                         "frac_of_hbm_peak": rate * 2 * BYTES_LIMB / 1e9 / peak,
                         "frac_of_integer_pipe_ceiling": rate * 0.236e-6}
        ntt["note"] = ("N = 2^15 limbs, in place, canonical in/out; the integer-pipe ceiling is 0.236 us per limb "
                       "(tools/pipe_bench.cu, DESIGN.md section 3), the HBM roofline 0.080 us")
        line["ntt"] = ntt
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        subprocess.run(["make", "-s", "-C", str(REPO / "oracle")], check=True)
        kd = tempfile.mkdtemp(prefix="hevm_cpu_keys_")
        one_step, ops = cpu_chain_rate(1, 1, [kd])
        one_step()
        reps = 3
        t = sum(one_step() for _ in range(reps))
        line["cpu_baseline"] = {"value": ops * reps / t, "unit": "ops/s", "cores": 1, "kind": "port",
                                "sample": f"1 ciphertext chain ({ops} HEVM ops, levels 13..1) x {reps} repetitions on one host core "
                                          f"(the reference's SEAL evaluator is single-threaded); host has {os.cpu_count()} cores"}
    if oracle_proc is not None:  # the MEASURED run() of the CPU port on the very ResNet-20 program
        try:
            o, _ = oracle_proc.communicate(timeout=600)
            cpu = json.loads(o.strip().splitlines()[-1])
            line["resnet20"]["cpu_port_s"] = cpu["run_s"]
            line["resnet20"]["speedup_vs_cpu_port"] = cpu["run_s"] / line["resnet20"]["run_latency_s"]
            line["resnet20"]["speedup_vs_cpu_port_first_run"] = cpu["run_s"] / line["resnet20"]["first_run_s"]
            details["resnet20"]["cpu_port"] = cpu
        except Exception as e:
            line["resnet20"]["cpu_port_s"] = None
            details["resnet20"]["cpu_port_error"] = repr(e)
            oracle_proc.kill()
    if world > 1 and not args.no_sharded:
        # BASELINE.json configs[4]: RNS-limb-sharded key switch at N = 2^16 over the N GPUs (NCCL all-gather of the digits
        # + broadcast of the rounded special limb); not part of `value`, reported beside it
        from dacapo_b200 import sharded
        sk = sharded.measure(lib, rank, world, logn=16, nprimes=30)
        line["sharded_keyswitch"] = sk
    if rank == 0:
        try:
            (REPO / "bench_details.json").write_text(json.dumps({**line, **details}, indent=1))
        except OSError:
            pass
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def _main_with_clean_stdout():
    """The driver reads ONE JSON line from stdout: anything a library prints there (NCCL's version banner, setDebug traces)
    goes to stderr instead; the real stdout is restored only around the final print (json lines go through _emit)."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    try:
        main()
    finally:
        sys.stdout.flush()
        os.dup2(_REAL_STDOUT, 1)


_REAL_STDOUT = None


def _emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())
    else:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    _main_with_clean_stdout()
