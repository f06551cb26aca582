#!/usr/bin/env python
"""Headline benchmark: fwd+bwd gate-applies/sec of a random brickwork circuit.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--qubits 32] [--depth 100] [--precision f32] [--fuse 1]

Workload (BASELINE.json configs[3]): n-qubit brickwork of
depth D -- layer l applies `add_q2_var_gate(i+1, i)` for i = l mod 2, l mod 2 + 2, ...
-- Haar-random 4x4 gates (NumPy default_rng(1234)), initial state |0..0>, loss
drivers `get_q2_dens_op_with_grad(i+1, i)` on every even i with a fixed random
Hermitian cotangent.  One STEP = one `Circuit.forward` + one `Circuit.backward`
through the reference-facing API (host NumPy gate lists in, host densities and
per-gate gradients out), i.e. the call a user of the reference makes.

Metric: gate-applies/s = (#gate instructions) / (time of forward + backward).
  value : timed on the device with CUDA events around the K steps.
  e2e   : wall clock around the same public-API calls (host marshalling, H2D of
          the gate matrices as kernel parameters, D2H of densities/gradients).
The state is generated on the device by the API itself (|0..0>, as in the
reference's Circuit::new); the host inputs of the path are the gate lists.

`--impl reference` drives the reference's own CUDA library (oracle/_ref, built
unmodified from /root/reference/src/primitives.cu) with the exact call sequence
of its Rust `Circuit` (oracle/ref_replay.py) on a depth-bounded sample of the
same workload; the reference has no CPU implementation of the path.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fwd+bwd gate-applies/sec"
UNIT = "gate-applies/s"


# ------------------------------------------------------------------ workload
def haar(rng, k):
    z = rng.normal(size=(k, k)) + 1j * rng.normal(size=(k, k))
    q, r = np.linalg.qr(z)
    return (q * (np.diag(r) / np.abs(np.diag(r)))).reshape(-1)


def brickwork_program(n, depth):
    """[(pos2, pos1)] of the gate instructions and of the trailing densities."""
    gates = []
    for layer in range(depth):
        for i in range(layer % 2, n - 1, 2):
            gates.append((i + 1, i))
    dens = [(i + 1, i) for i in range(0, n - 1, 2)]
    return gates, dens


def build_brickwork(circuit, n, depth):
    gates, dens = brickwork_program(n, depth)
    for p2, p1 in gates:
        circuit.add_q2_var_gate(p2, p1)
    for p2, p1 in dens:
        circuit.get_q2_dens_op_with_grad(p2, p1)
    return len(gates), len(dens)


def brickwork_inputs(n, depth, dtype):
    rng = np.random.default_rng(1234)
    gates, dens = brickwork_program(n, depth)
    var = [haar(rng, 4).astype(dtype) for _ in gates]
    cts = []
    for _ in dens:
        a = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
        cts.append(((a + a.conj().T) / 2).astype(dtype))
    return var, cts


def vqse_program(n, layers):
    """example_vqse_ising.py:66-79: per layer n ring ZZ diagonal gates + n X rotations; n ring densities."""
    prog = []
    for _ in range(layers):
        for i in range(n - 1):
            prog.append(("diag", i, i + 1))
        prog.append(("diag", 0, n - 1))
        for i in range(n):
            prog.append(("q1", i, -1))
    dens = [(i, i + 1) for i in range(n - 1)] + [(0, n - 1)]
    return prog, dens


def build_vqse(circuit, n, layers):
    prog, dens = vqse_program(n, layers)
    for kind, a, b in prog:
        if kind == "diag":
            circuit.add_q2_var_gate_diag(a, b)
        else:
            circuit.add_q1_var_gate(a)
    for a, b in dens:
        circuit.get_q2_dens_op_with_grad(a, b)
    return len(prog), len(dens)


def vqse_inputs(n, layers, dtype):
    """Gates of example_vqse_ising.py:15-28 from 2*layers N(0,1) parameters (seed 42); cotangent = h^T of the
    transverse-field Ising two-site term (:86-93)."""
    rng = np.random.default_rng(42)
    params = rng.normal(size=2 * layers)
    var = []
    for l in range(layers):
        g, b = params[2 * l], params[2 * l + 1]
        zz = np.array([np.exp(-1j * g), np.exp(1j * g), np.exp(1j * g), np.exp(-1j * g)], dtype=dtype)
        x = np.array([np.cos(b), -1j * np.sin(b), -1j * np.sin(b), np.cos(b)], dtype=dtype)
        var += n * [zz] + n * [x]
    sz = np.diag([1.0, -1.0]); sx = np.array([[0, 1.0], [1.0, 0]]); eye = np.eye(2)
    k = lambda a, b: np.tensordot(a, b, axes=0).transpose(0, 2, 1, 3).reshape(4, 4)  # noqa: E731
    h = (-k(sz, sz) - 0.5 * (k(sx, eye) + k(eye, sx))).astype(dtype)
    _, dens = vqse_program(n, layers)
    return var, [h.T.copy() for _ in dens]


# -------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                   r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------ CPU baseline
def _cpu_baseline_numpy(budget_s):
    """The oracle (NumPy port of the path) on one core: fwd+bwd of brickwork gates at 22 qubits."""
    from oracle import statevector as sv
    n = 22
    rng = np.random.default_rng(1234)
    dtype = np.complex64
    psi = sv.standard_state(n, dtype)
    gates, _ = brickwork_program(n, 8)    # the forward loop stops at budget / 3
    mats = [haar(rng, 4).astype(dtype) for _ in gates]
    t0 = time.perf_counter()
    done = 0
    for (p2, p1), g in zip(gates, mats):   # forward
        psi = sv.q2gate_fast(psi, g, p2, p1)
        done += 1
        if time.perf_counter() - t0 > budget_s / 3:
            break
    bwd = sv.conj_and_double(psi)
    for (p2, p1), g in list(zip(gates, mats))[:done][::-1]:   # backward: un-compute, grad, adjoint
        psi = sv.q2gate_fast(psi, sv.q2_conj_tr(g), p2, p1)
        sv.q2grad(psi, bwd, p2, p1)
        bwd = sv.q2gate_fast(bwd, sv.q2_tr(g), p2, p1)
    dt = time.perf_counter() - t0
    return n, done, dt


def _torch_apply_adjacent(x, g, lo, n):
    """q2 gate (4x4 torch tensor, reference index convention) on the adjacent pair (lo + 1, lo) of a torch state."""
    import torch
    return torch.matmul(g, x.view(1 << (n - lo - 2), 4, 1 << lo)).reshape(-1)


def _cpu_baseline_torch(budget_s):
    """The same algorithm (apply / un-compute with U^dagger / gradient outer product / pull back with U^T) as
    batched 4x4 matmuls on torch CPU tensors with all host threads: fwd+bwd of brickwork gates at 24 qubits."""
    import torch
    n = 24
    rng = np.random.default_rng(1234)
    gates, _ = brickwork_program(n, 40)   # more than the time budget admits: the forward loop stops at budget / 4
    mats = [torch.tensor(haar(rng, 4).reshape(4, 4), dtype=torch.complex64) for _ in gates]
    psi = torch.zeros(1 << n, dtype=torch.complex64)
    psi[0] = 1

    def apply(x, g, lo):
        return _torch_apply_adjacent(x, g, lo, n)

    t0 = time.perf_counter()
    done = 0
    for (p2, p1), g in zip(gates, mats):
        psi = apply(psi, g, p1)
        done += 1
        if time.perf_counter() - t0 > budget_s / 4:
            break
    bwd = 2 * psi.conj()
    for (p2, p1), g in list(zip(gates, mats))[:done][::-1]:
        psi = apply(psi, g.conj().T.contiguous(), p1)
        torch.einsum("apc,aqc->pq", bwd.view(1 << (n - p1 - 2), 4, 1 << p1), psi.view(1 << (n - p1 - 2), 4, 1 << p1))
        bwd = apply(bwd, g.T.contiguous(), p1)
    dt = time.perf_counter() - t0
    return n, done, dt, torch.get_num_threads()


def cpu_baseline(n_target, budget_s=12.0):
    """CPU restatements of the path timed on the host (the reference has no CPU implementation): the NumPy oracle
    on one core and a torch port on all host threads, each on a reduced register (the 2^32-amplitude state does not
    fit a bounded CPU run), scaled to the target size by the amplitude ratio.  The faster one is reported."""
    n1, done1, dt1 = _cpu_baseline_numpy(budget_s)
    r1 = done1 / dt1 * 2.0 ** (n1 - n_target)
    note1 = (f"oracle (NumPy, 1 core) fwd+bwd of {done1} brickwork gates at {n1} qubits complex64 in {dt1:.1f}s "
             f"= {done1 / dt1:.3g} gate-applies/s")
    try:
        n2, done2, dt2, threads = _cpu_baseline_torch(budget_s)
        r2 = done2 / dt2 * 2.0 ** (n2 - n_target)
        note2 = (f"torch-CPU port ({threads} threads of {os.cpu_count()} cores) fwd+bwd of {done2} brickwork gates at "
                 f"{n2} qubits complex64 in {dt2:.1f}s = {done2 / dt2:.3g} gate-applies/s")
    except Exception as e:  # noqa: BLE001
        r2, threads, note2 = 0.0, 1, f"torch-CPU port failed: {e}"
    best, cores = (r2, threads) if r2 > r1 else (r1, 1)
    return {"value": best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{note2}; {note1}; both scaled by the amplitude ratio to {n_target} qubits, the faster reported"}


# ------------------------------------------------------------------- arms
def dist_setup(n_gpus):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def run_ours(args):
    import torch
    rank, world, local = dist_setup(args.gpus)
    pkg = importlib.import_module("differentiable-quantum-circuit-cuda_b200")
    dtype = np.complex64 if args.precision == "f32" else np.complex128
    g = world.bit_length() - 1
    n_total = args.qubits + g          # weak scaling: 32 local qubits per GPU
    if world > 1:
        from importlib import import_module
        sharded = import_module("differentiable-quantum-circuit-cuda_b200.sharded")
        circ = sharded.ShardedCircuit(n_total, precision=args.precision)
    elif args.fusion:
        circ = pkg.FusedCircuit(n_total, precision=args.precision)
    else:
        circ = pkg.Circuit(n_total, precision=args.precision)
    circ.set_option("fuse", args.fuse)
    circ.set_option("profile", 1)
    if args.tile_bits:
        circ.set_option("tile_bits", args.tile_bits)
    if args.low_bits:
        circ.set_option("low_bits", args.low_bits)
    if args.max_tile_gates:
        circ.set_option("max_tile_gates", args.max_tile_gates)
    if args.tile_debug:
        circ.set_option("tile_debug", args.tile_debug)
    if world > 1:
        circ.set_option("peer", args.peer)
    if args.precision == "f32":
        circ.set_option("soa", args.soa)
    if args.rb_policy >= 0:
        circ.set_option("rb_policy", args.rb_policy)
    if args.batch_dens >= 0:
        circ.set_option("batch_dens", args.batch_dens)
    if args.tile_strategy >= 0:
        circ.set_option("tile_strategy", args.tile_strategy)
    if args.stagger >= 0:
        circ.set_option("stagger", args.stagger)
    if args.workload == "vqse":
        n_gates, n_dens = build_vqse(circ, n_total, args.depth)
        var, cts = vqse_inputs(n_total, args.depth, dtype)
        if world == 1:
            circ.set_state_from_vector((np.ones(1 << n_total) / np.sqrt(float(1 << n_total))).astype(dtype))
    else:
        n_gates, n_dens = build_brickwork(circ, n_total, args.depth)
        var, cts = brickwork_inputs(n_total, args.depth, dtype)
    cts_conj = [c.conj() for c in cts]
    h2d = sum(v.nbytes for v in var) * 2 + sum(c.nbytes for c in cts)   # gates go in twice (fwd, bwd)
    d2h = n_dens * 16 * var[0].itemsize + sum(v.nbytes for v in var)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        dens = circ.forward([], var)
        grads = circ.backward(cts_conj, [], var)
        return dens, grads

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches, prof = 0, {}
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        dens = circ.forward([], var)
        s_f, p_f = circ.last_stats(), circ.last_profile()
        grads = circ.backward(cts_conj, [], var)
        s_b, p_b = circ.last_stats(), circ.last_profile()
        launches += s_f["kernel_launches"] + s_b["kernel_launches"]
        for p in (p_f, p_b):
            for k, v in p.items():
                e = prof.setdefault(k, {"launches": 0, "ms": 0.0, "algorithmic_bytes": 0})
                for kk in e:
                    e[kk] += v[kk]
    ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([dev_ms, wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall = float(t[0]), float(t[1])
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    # dominant kernel = the category with the most device time
    dom = max(prof.items(), key=lambda kv: kv[1]["ms"]) if prof else (None, None)
    roofline = None
    if dom[0]:
        e = dom[1]
        achieved = e["algorithmic_bytes"] / (e["ms"] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom[0], "achieved": round(achieved, 1), "peak": peak,
                    "peak_source": peak_src, "unit": "GB/s", "frac": round(achieved / peak, 4),
                    "traffic": None, "launches": e["launches"],
                    "avg_launch_ms": round(e["ms"] / e["launches"], 4),
                    "algorithmic_bytes_per_launch": e["algorithmic_bytes"] // e["launches"],
                    "share_of_step": round(e["ms"] / dev_ms, 4)}
    # DRAM traffic of one launch of the dominant kernel: the ncu capture under profiles/ (one reverse pass
    # reads + writes state and adjoint once = 4*S whatever the number of fused gates), scaled by the shard size
    if roofline and roofline["kernel"] == "tile_bwd":
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
            if tr["precision"] == args.precision:
                scale = 2.0 ** (args.qubits - tr["qubits"])
                roofline["traffic"] = int((tr["dram_bytes_read"] + tr["dram_bytes_write"]) * scale)
                roofline["traffic_source"] = ("ncu dram__bytes_read+write of one k_tile_bwd_soa launch at %d q "
                                              "(profiles/r1_traffic_32q_bwd.csv)" % tr["qubits"]) + \
                                             ("" if scale == 1.0 else ", scaled by the shard size")
                roofline["traffic_note"] = ("a launch applies %.1f gates on average: algorithmic bytes = 4*S per "
                                            "gate, DRAM traffic = 4*S per launch (no re-reads)"
                                            % (roofline["algorithmic_bytes_per_launch"] / (4.0 * (int(np.dtype(dtype).itemsize) << args.qubits))))
        except Exception:  # noqa: BLE001
            pass
    total_alg = sum(e["algorithmic_bytes"] for e in prof.values())
    # Work unit = one gate applied to one 2^local_qubits-amplitude shard.  A sharded run applies every
    # gate to `world` shards, so the whole-job aggregate is world * gates / time (weak scaling: per-GPU
    # work per gate is fixed; ideal aggregate grows linearly with the number of GPUs).
    value = world * n_gates * args.steps / (dev_ms * 1e-3)
    out = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(dev_ms / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": f"{args.workload}-{n_total}q-depth{args.depth}-{args.precision}", "qubits": n_total,
                   "local_qubits": args.qubits, "depth": args.depth, "gates": n_gates, "densities": n_dens,
                   "state_bytes_per_gpu": int(np.dtype(dtype).itemsize) << args.qubits,
                   "unit_note": "one gate-apply = one gate applied (fwd+bwd) to one 2^local_qubits-amplitude shard; "
                                "an N-GPU run applies each gate to N shards",
                   "full_state_gate_applies_per_s": round(n_gates * args.steps / (dev_ms * 1e-3), 3),
                   "l2": "inputs (state + adjoint) far larger than L2; no flush needed",
                   "exchange": (("peer-memory swap kernel (NVLink, CUDA IPC)" if circ.peer_exchange else
                                 "NCCL send/recv + pack/unpack") if world > 1 else None),
                   "gate_fusion": bool(args.fusion and world == 1),
                   "executor": ["one pass per gate", "tiled multi-gate passes",
                                "tiled multi-gate passes, forward kernel (register-blocked or per gate) by gate mix"][args.fuse]
                               + (", pair-lane smem layout" if args.precision == "f32" and args.soa and args.fuse else "")},
        "effective_hbm_gbs": round(total_alg * world / (dev_ms * 1e-3) / 1e9, 1),
        "effective_hbm_frac": round(total_alg / (dev_ms * 1e-3) / 1e9 / peak, 4),
        "roofline": roofline,
        "profile_ms": {k: round(v["ms"], 2) for k, v in sorted(prof.items())},
        "e2e": {"value": round(world * n_gates * args.steps / wall, 3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "check": {"trace_density0": float(np.trace(dens[0]).real),
                  "grad_norm": float(np.sqrt(sum(np.vdot(x, x).real for x in grads)))},
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(n_total)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_replay as rr
    import torch
    torch.cuda.set_device(0)
    n = args.qubits
    dtype = np.complex64 if args.precision == "f32" else np.complex128
    if not rr.ref_available(args.precision, big=n > 30):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (make -C oracle)"}))
        return
    depth = args.ref_depth
    circ = rr.RefCircuit(n, args.precision)
    n_gates, n_dens = build_brickwork(circ, n, depth)
    var, cts = brickwork_inputs(n, depth, dtype)
    cts_conj = [c.conj() for c in cts]

    def step():
        circ.forward([], var)
        return circ.backward(cts_conj, [], var)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    value = n_gates * args.steps / (dev_ms * 1e-3)
    sample = (f"reference CUDA library ({os.path.basename(circ.lib.path)}) replaying src/circuit.rs on "
              f"brickwork-{n}q depth {depth} of 100 ({n_gates} gates + {n_dens} densities per step), one B200")
    out = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dev_ms / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision,
        "data": "synthetic",
        "config": {"workload": f"brickwork-{n}q-depth{args.depth}-{args.precision}", "qubits": n,
                   "depth_sampled": depth, "gates": n_gates},
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": 0, "kind": "reference",
                         "sample": sample + " -- the reference has no CPU implementation; this is its own GPU code"},
        "e2e": {"value": round(n_gates * args.steps / wall, 3), "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--qubits", type=int, default=32, help="LOCAL qubits per GPU (total = qubits + log2(gpus))")
    ap.add_argument("--depth", type=int, default=100, help="brickwork layers / VQSE (ZZ, X) layer pairs")
    ap.add_argument("--workload", default="brickwork", choices=["brickwork", "vqse"])
    ap.add_argument("--ref-depth", type=int, default=2, help="depth of the bounded sample the reference arm runs")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--fuse", type=int, default=2, help="0: one pass per gate, 1: tiled passes, 2: register-blocked tiled passes")
    ap.add_argument("--tile-bits", type=int, default=0, help="T of the tiled passes (0: library default)")
    ap.add_argument("--low-bits", type=int, default=0, help="L lowest positions forced into every tile (0: default)")
    ap.add_argument("--max-tile-gates", type=int, default=0)
    ap.add_argument("--peer", type=int, default=1, help="sharded: 1 peer-memory swap kernel, 0 NCCL send/recv")
    ap.add_argument("--soa", type=int, default=1, help="f32 tile kernels: 1 pair-lane smem layout, 0 interleaved layout")
    ap.add_argument("--fusion", type=int, default=0, help="1: FusedCircuit (host-side gate fusion + gradient chain rule)")
    ap.add_argument("--rb-policy", type=int, default=-1, help="fuse=2 forward: 0 never register-block, 1 always, 2 by gate mix (default)")
    ap.add_argument("--batch-dens", type=int, default=-1, help="0: one sweep per density / seed")
    ap.add_argument("--tile-strategy", type=int, default=-1, help="scheduler tiling: 2 window growth with look-ahead (default), 1 window growth, 0 first-fit")
    ap.add_argument("--stagger", type=int, default=-1, help="CTA start skew, percent of the library default (0: off)")
    ap.add_argument("--tile-debug", type=int, default=0, help="profiling aid (1: no HBM traffic, 2: no gates); invalid results")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
