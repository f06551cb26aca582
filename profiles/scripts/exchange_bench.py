"""Half-shard exchange bandwidth vs the local position that is swapped with the global qubit.
    torchrun --nproc-per-node 2 profiles/scripts/exchange_bench.py [local_qubits]
A circuit is built whose only possible remap victim is local qubit v (every other local qubit has a
pending gate behind the blocked global-qubit gate), so the plan holds exactly one swap (gbit 0, lpos v);
the "exchange" profile category then times one exchange of 2^(n_loc-1) amplitudes per direction."""
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
    n_loc = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank()
    importlib.import_module("differentiable-quantum-circuit-cuda_b200")
    sharded = importlib.import_module("differentiable-quantum-circuit-cuda_b200.sharded")
    n = n_loc + 1
    rng = np.random.default_rng(0)
    z = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    g = np.linalg.qr(z)[0].astype(np.complex64).reshape(-1)
    rows = []
    for peer in (1, 0):
        for v in (0, 1, 2, 3, 4, 5, 6, 8, 12, 20, n_loc - 2):
            c = sharded.ShardedCircuit(n, precision="f32")
            c.set_option("fuse", 0)
            c.set_option("profile", 1)
            c.set_option("peer", peer)
            c.add_q2_const_gate(n - 1, n_loc - 1)
            cnt = 1
            for q in range(n_loc - 1):
                if q != v:
                    c.add_q2_const_gate(q, n_loc - 1)
                    cnt += 1
            c.get_q1_dens_op(0)
            for it in range(2):
                c.run([g] * cnt, [])
                p = c.last_profile()
            ms = p["exchange"]["ms"] / p["exchange"]["launches"]
            gbs = (8 << (n_loc - 1)) / (ms * 1e-3) / 1e9
            rows.append({"peer_kernel": bool(c.peer_exchange), "lpos": v, "launches": p["exchange"]["launches"],
                         "ms": round(ms, 3), "GBps_per_direction": round(gbs, 1)})
            del c
    if rank == 0:
        for r in rows:
            print(json.dumps(r))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
