"""GPU, world_size > 1: ShardedCircuit (half-shard exchanges between GPUs of one box,
as the NVLink peer-memory swap kernel (peer=1) or NCCL send/recv (peer=0)) must reproduce the single-GPU executor and the oracle.  Needs >= 2
visible GPUs (`gpurun --gpus 2`); skipped on a single-GPU box."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, case, n, precision, fuse, peer, q, tc=0):
    """A failing rank reports through the queue: the parent must not sit out its timeout on a multi-GPU box."""
    try:
        _worker_body(rank, world, port, case, n, precision, fuse, peer, q, tc)
    except BaseException as e:  # noqa: BLE001
        import traceback
        q.put((rank, "error", f"{type(e).__name__}: {e}\n{traceback.format_exc()}", 0.0))
        raise


def _worker_body(rank, world, port, case, n, precision, fuse, peer, q, tc=0):
    import importlib
    import torch
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    importlib.import_module("differentiable-quantum-circuit-cuda_b200")
    sharded = importlib.import_module("differentiable-quantum-circuit-cuda_b200.sharded")
    from test_scheduler import make_case, reference_results
    dtype = np.complex64 if precision == "f32" else np.complex128
    o, const, var = make_case(case, n, np.random.default_rng(5))
    dens_ref, cts_conj, grads_ref = reference_results(o, const, var)
    c = sharded.ShardedCircuit(n, precision=precision)
    c.set_option("fuse", fuse)
    if tc:
        c.set_option("tc", 1)          # tensor-core 6-qubit blocks on every shard (default only for shards >= 2^26)
        c.set_option("profile", 1)
    else:
        c.set_option("tile_bits", 11)
    c.set_option("peer", peer)
    if peer:
        assert c.peer_exchange, "CUDA IPC peer mapping failed on an NVLink box"
    for inst in o.instructions:
        c._add(*inst)
    cg, vg = [g.astype(dtype) for g in const], [g.astype(dtype) for g in var]
    dens = c.forward(cg, vg)
    grads = c.backward([x.astype(dtype) for x in cts_conj], cg, vg)
    if tc and case != "autodiff":   # (autodiff: the windows hold NonU gates and keep the FP32 tile kernels in the reverse pass)
        assert c.last_profile().get("tc_bwd", {}).get("launches", 0) > 0, "the tensor-core blocks must be the ones that ran"
    if tc and world >= 4:
        # merged multi-qubit exchanges (k_peer_multiswap) against one exchange per swapped qubit: pure data movement,
        # so the results agree to the last bits, with fewer exchange launches
        ex1 = c.last_profile().get("exchange", {}).get("launches", 0)
        c0 = sharded.ShardedCircuit(n, precision=precision)
        for k, v in (("fuse", fuse), ("tc", 1), ("profile", 1), ("peer", peer), ("multi_swap", 0)):
            c0.set_option(k, v)
        for inst in o.instructions:
            c0._add(*inst)
        dens0 = c0.forward(cg, vg)
        grads0 = c0.backward([x.astype(dtype) for x in cts_conj], cg, vg)
        ex0 = c0.last_profile().get("exchange", {}).get("launches", 0)
        sc = max(np.abs(x).max() for x in grads0)
        assert max(np.abs(a - b).max() for a, b in zip(dens, dens0)) < 1e-6, "merged exchange changed the densities"
        assert max(np.abs(a - b).max() for a, b in zip(grads, grads0)) / sc < 1e-6, "merged exchange changed the gradients"
        assert ex1 <= ex0, (ex1, ex0)
        if case == "brickwork":
            assert ex1 < ex0, ("no run of swaps was merged", ex1, ex0)
        del c0
    err_d = max(np.abs(a - b).max() for a, b in zip(dens, dens_ref))
    scale = max(np.abs(x).max() for x in grads_ref)
    err_g = max(np.abs(a - b).max() for a, b in zip(grads, grads_ref)) / scale
    shard = c.get_cpu_state_copy()
    init = np.zeros_like(shard)
    if rank == 0:
        init[0] = 1
    q.put((rank, float(err_d), float(err_g), float(np.abs(shard - init).max())))
    dist.destroy_process_group()


# world 2: every executor / exchange combination; world 4 and 8 (the rank counts of the SCALE runs): the default
# executor with both exchange implementations
_COMBOS = [(2, f, p) for f, p in [(0, 1), (1, 0), (2, 1), (2, 0)]] + [(w, 2, p) for w in (4, 8) for p in (1, 0)]


@pytest.mark.parametrize("case", ["brickwork", "autodiff", "vqse"])
@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("world,fuse,peer", _COMBOS)
def test_sharded_matches_oracle(case, precision, world, fuse, peer):
    import torch.multiprocessing as mp
    if _ngpus() < world:
        pytest.skip(f"needs >= {world} GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 15
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, n, precision, fuse, peer, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    assert not any(r[1] == "error" for r in results), [r[2] for r in results if r[1] == "error"]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    tol = 1e-5 if precision == "f32" else 1e-12   # north_star tolerances
    for rank, err_d, err_g, err_s in results:
        assert err_d < tol and err_g < tol and err_s < tol * 10, (rank, err_d, err_g, err_s)


@pytest.mark.parametrize("case", ["brickwork", "autodiff", "vqse"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_tensor_core_blocks_match_oracle(case, world):
    """The f32 default of the measured sizes (tensor-core blocks, fused reverse step) on sharded registers: 2^14-amplitude
    shards (the smallest a block pass accepts) with half-shard exchanges between the blocks."""
    import torch.multiprocessing as mp
    if _ngpus() < world:
        pytest.skip(f"needs >= {world} GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 14 + world.bit_length() - 1
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, n, "f32", 2, 1, q, 1)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    assert not any(r[1] == "error" for r in results), [r[2] for r in results if r[1] == "error"]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, err_d, err_g, err_s in results:
        assert err_d < 1e-5 and err_g < 1e-5 and err_s < 1e-4, (rank, err_d, err_g, err_s)
