"""Host-side reader / writer of the state checkpoint format (include/qdc_circuit.h,
`qdc_circuit_save_state`): one file per rank, 128-byte header + the rank's shard
as raw little-endian interleaved (re, im) pairs in PHYSICAL index order.

Pure NumPy (no GPU, no library): used to build initial-state files, to inspect
dumps, and -- for registers that fit host memory -- to assemble the shards of a
sharded run into the canonical vector (index = sum_k bit_k 2^k, qubit k = bit k,
the reference's layout, /root/reference/src/qdc/circuit.py:29-30).
"""
from __future__ import annotations

import struct
from typing import Dict, Sequence, Tuple

import numpy as np

MAGIC = b"QDCSTAT1"
HEADER_BYTES = 128
_FMT = "<8s6I64s32s"


def _dtype(real_bytes: int):
    if real_bytes == 4:
        return np.dtype("<c8")
    if real_bytes == 8:
        return np.dtype("<c16")
    raise ValueError(f"unsupported real size {real_bytes}")


def pack_header(real_bytes: int, n: int, n_loc: int, rank: int, world: int, qubit_map: Sequence[int]) -> bytes:
    if n > 64:
        raise ValueError("state files hold at most 64 qubits")
    m = bytearray(b"\xff" * 64)
    for q, p in enumerate(qubit_map):
        m[q] = p
    return struct.pack(_FMT, MAGIC, real_bytes, n, n_loc, rank, world, 0, bytes(m), b"\0" * 32)


def read_header(path) -> Dict:
    with open(path, "rb") as f:
        raw = f.read(HEADER_BYTES)
    if len(raw) != HEADER_BYTES:
        raise ValueError(f"{path} is not a state file (short header)")
    magic, real_bytes, n, n_loc, rank, world, _, m, _ = struct.unpack(_FMT, raw)
    if magic != MAGIC:
        raise ValueError(f"{path} is not a state file (bad magic)")
    return {"real_bytes": real_bytes, "n": n, "n_loc": n_loc, "rank": rank, "world": world,
            "map": [int(x) for x in m[:n]]}


def read_shard(path) -> Tuple[Dict, np.ndarray]:
    """(header, shard in physical order).  The payload is memory-mapped."""
    h = read_header(path)
    data = np.memmap(path, dtype=_dtype(h["real_bytes"]), mode="r", offset=HEADER_BYTES, shape=(1 << h["n_loc"],))
    return h, data


def write_shard(path, shard: np.ndarray, n: int, rank: int = 0, world: int = 1, qubit_map=None):
    """Write one rank's shard (physical order).  `qubit_map` defaults to the identity."""
    shard = np.ascontiguousarray(shard)
    if shard.dtype not in (np.complex64, np.complex128):
        raise TypeError("shard must be complex64 or complex128")
    g = world.bit_length() - 1
    if world != 1 << g or shard.size != 1 << (n - g):
        raise ValueError("shard size does not match n and world")
    with open(path, "wb") as f:
        f.write(pack_header(shard.dtype.itemsize // 2, n, n - g, rank, world,
                            list(range(n)) if qubit_map is None else qubit_map))
        shard.astype(shard.dtype.newbyteorder("<"), copy=False).tofile(f)


def assemble(paths: Sequence) -> np.ndarray:
    """Canonical full state from the per-rank files of one dump (host memory: 2^n entries)."""
    shards = sorted((read_shard(p) for p in paths), key=lambda hs: hs[0]["rank"])
    h0 = shards[0][0]
    n, n_loc, world = h0["n"], h0["n_loc"], h0["world"]
    if len(shards) != world or [h["rank"] for h, _ in shards] != list(range(world)):
        raise ValueError("need exactly one file per rank")
    for h, _ in shards:
        if (h["n"], h["n_loc"], h["world"], h["real_bytes"], h["map"]) != \
                (n, n_loc, world, h0["real_bytes"], h0["map"]):
            raise ValueError("files belong to different dumps")
    phys = np.concatenate([np.asarray(d) for _, d in shards])  # physical index = (rank << n_loc) | local index
    qmap = h0["map"]
    if qmap == list(range(n)):
        return phys
    # tensor axis of physical position p is n-1-p; logical qubit q must end up on axis n-1-q
    t = phys.reshape([2] * n)
    axes = [n - 1 - qmap[n - 1 - ax] for ax in range(n)]
    return np.ascontiguousarray(t.transpose(axes)).reshape(-1)
