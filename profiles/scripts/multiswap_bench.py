"""Merged multi-qubit exchanges (k_peer_multiswap) against one exchange per swapped qubit on the same sharded circuit.
    torchrun --nproc-per-node 4 profiles/scripts/multiswap_bench.py [local_qubits=28] [depth=24]
Prints the exchange time, launches and the rate per direction of both settings, and checks that the results agree."""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    n_loc = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    depth = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    g = world.bit_length() - 1
    importlib.import_module("differentiable-quantum-circuit-cuda_b200")
    sharded = importlib.import_module("differentiable-quantum-circuit-cuda_b200.sharded")
    bench = importlib.import_module("bench")
    n = n_loc + g
    var, cts = bench.brickwork_inputs(n, depth, np.complex64)
    out = {}
    for multi in (1, 0):
        c = sharded.ShardedCircuit(n, precision="f32")
        c.set_option("profile", 1)
        c.set_option("multi_swap", multi)
        bench.build_brickwork(c, n, depth)
        for _ in range(2):
            dens = c.forward([], var)
            pf = c.last_profile()
            grads = c.backward([x.conj() for x in cts], [], var)
            pb = c.last_profile()
        ms = pf["exchange"]["ms"] + pb["exchange"]["ms"]
        launches = pf["exchange"]["launches"] + pb["exchange"]["launches"]
        out[multi] = (dens, grads, ms, launches)
        del c
    if rank == 0:
        shard = 8 << n_loc
        for multi in (1, 0):
            _, _, ms, launches = out[multi]
            print(json.dumps({"world": world, "local_qubits": n_loc, "depth": depth, "multi_swap": multi, "exchange_ms": round(ms, 2),
                              "exchange_launches": launches}))
        d = max(float(np.abs(a - b).max()) for a, b in zip(out[1][0], out[0][0]))
        gd = max(float(np.abs(a - b).max()) for a, b in zip(out[1][1], out[0][1]))
        print(json.dumps({"max_abs_diff_densities": d, "max_abs_diff_gradients": gd, "shard_bytes": shard}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
