"""Independent restatement of the reference's stored values.  TEST INFRASTRUCTURE ONLY (see oracle/README.md).

The reference persists with ``bincode::serde::encode_to_vec(x, bincode::config::legacy())``
(/root/reference/src/database/index/lsh.rs:94, :102, :108-118; src/database/core.rs:95, :186).  bincode 2's
``legacy()`` configuration = little endian, fixed-width integers, ``u64`` sequence lengths, ``u32`` enum variant
indices; through serde a ``Uuid`` is a byte string (``u64`` length 16 + 16 bytes), a ``Box`` is transparent, a unit
struct is empty and ``Embedding<N>`` (serde_with ``[_; N]``, src/lib.rs:16-18) is a tuple of N f32 with no length.

bincode / uuid / serde_with are un-vendored dependencies (Cargo.toml:56-59) -- PARITY UNPINNED: the layout is restated
from their published formats and pinned by the hand-written byte strings in tests/test_interchange.py.

Types follow the reference's definitions one to one, as a recursive Python structure:
    Node::Inner(InnerNode { hyperplane: Hyperplane { coefficients, constant }, left_node, right_node })   lsh.rs:45-57
        -> ("inner", coefficients: np.float32[N], constant: float, left, right)
    Node::Leaf(LeafNode(Vec<Uuid>))                                                                      lsh.rs:59-60
        -> ("leaf", [uuid.UUID, ...])
"""
from __future__ import annotations

import struct
import sys
import uuid
from typing import List, Tuple

import numpy as np


def encode_uuid(u: uuid.UUID) -> bytes:
    return struct.pack("<Q", 16) + u.bytes


def encode_embedding(x) -> bytes:
    """lsh.rs:91-97: the value stored under a vector id."""
    return np.asarray(x, dtype="<f4").tobytes()


def decode_embedding(b: bytes, dim: int) -> np.ndarray:
    if len(b) != 4 * dim:
        raise ValueError("embedding blob has the wrong size")
    return np.frombuffer(b, dtype="<f4").astype(np.float32)


def encode_node(node) -> bytes:
    """lsh.rs:99-105: the value stored under a tree id (recursive, mirrors serde's derive order)."""
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 10000))
    if node[0] == "inner":
        _, coef, cst, left, right = node
        return struct.pack("<I", 0) + encode_embedding(coef) + struct.pack("<f", cst) + encode_node(left) + encode_node(right)
    _, ids = node
    return struct.pack("<IQ", 1, len(ids)) + b"".join(encode_uuid(u) for u in ids)


def decode_node(b: bytes, dim: int):
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 10000))
    node, at = _decode_node(memoryview(b), 0, dim)
    if at != len(b):
        raise ValueError("trailing bytes after the tree")
    return node


def _decode_node(b: memoryview, at: int, dim: int):
    (tag,) = struct.unpack_from("<I", b, at)
    at += 4
    if tag == 0:
        coef = np.frombuffer(b[at:at + 4 * dim], dtype="<f4").astype(np.float32)
        if coef.size != dim:
            raise ValueError("truncated hyperplane")
        at += 4 * dim
        (cst,) = struct.unpack_from("<f", b, at)
        at += 4
        left, at = _decode_node(b, at, dim)
        right, at = _decode_node(b, at, dim)
        return ("inner", coef, cst, left, right), at
    if tag == 1:
        (n,) = struct.unpack_from("<Q", b, at)
        at += 8
        ids = []
        for _ in range(n):
            (ln,) = struct.unpack_from("<Q", b, at)
            if ln != 16:
                raise ValueError("Uuid length is not 16")
            ids.append(uuid.UUID(bytes=bytes(b[at + 8:at + 24])))
            if len(b) < at + 24:
                raise ValueError("truncated Uuid")
            at += 24
        return ("leaf", ids), at
    raise ValueError(f"Node variant {tag}")


def encode_database_inner(db_uuid: uuid.UUID, max_node_size: int, num_trees: int, metric_power=None) -> bytes:
    """core.rs:19-29: uuid, model (unit struct), metric (unit struct, or {power: i32} for Minkowski / p-norm,
    distance.rs:162-165, :179-182), index_options {max_node_size: usize, num_trees: usize} (lsh.rs:124-129)."""
    out = encode_uuid(db_uuid)
    if metric_power is not None:
        out += struct.pack("<i", metric_power)
    return out + struct.pack("<QQ", max_node_size, num_trees)


# ---- flat forest (zb_index_export_forest numbering: preorder, trees in order) <-> recursive nodes ----
def forest_to_nodes(forest, ids_by_ordinal: List[uuid.UUID], live=None) -> list:
    """One recursive node per tree from the flat arrays; members become ids (only live ones if `live` is given)."""
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 10000))

    def rec(i):
        plane, left, right, leaf = (int(v) for v in forest.nodes[i])
        if plane >= 0:
            return ("inner", forest.coef.reshape(-1, forest.coef.shape[-1])[plane], float(forest.cst[plane]), rec(left), rec(right))
        m = forest.members[forest.leaf_off[leaf]:forest.leaf_off[leaf + 1]]
        return ("leaf", [ids_by_ordinal[int(o)] for o in m if live is None or live[int(o)]])

    return [rec(int(r)) for r in forest.roots]


def nodes_equal(a, b) -> bool:
    if a[0] != b[0]:
        return False
    if a[0] == "leaf":
        return a[1] == b[1]
    return (np.array_equal(np.asarray(a[1], np.float32).view(np.uint32), np.asarray(b[1], np.float32).view(np.uint32))
            and struct.pack("<f", a[2]) == struct.pack("<f", b[2]) and nodes_equal(a[3], b[3]) and nodes_equal(a[4], b[4]))
