"""CPU pins for the two-trees-per-warp kernel (alpharat_b200/csrc/mcts_half.cuh).

The kernel replaces compute_fpu's sum over the visited outcomes (search.rs:120-128) and the stored priors by table
lookups that are only valid because SmartUniform priors are all 1 / n (tree.rs:69-84).  These tests restate the two
facts the tables rest on in IEEE f32 (numpy), next to the oracle-backed GPU parity tests that exercise the kernel
itself; they also pin the shared-memory layout arithmetic used by the host to size the launches."""

from __future__ import annotations

import itertools

import numpy as np


def _visited_mass(n: int, visited: tuple[bool, ...]) -> np.float32:
    """compute_fpu's accumulation: outcomes in order, `mass += prior(i)` for the visited ones."""
    p = np.float32(1.0) / np.float32(n)
    mass = np.float32(0.0)
    for v in visited:
        if v:
            mass = np.float32(mass + p)
    return mass


def _table_entry(n: int, k: int) -> np.float32:
    """fpu_tab_entry(n, k) of mcts_half.cuh: k additions of the same f32, then sqrt."""
    p = np.float32(1.0) / np.float32(n)
    mass = np.float32(0.0)
    for _ in range(k):
        mass = np.float32(mass + p)
    return np.sqrt(mass, dtype=np.float32)


def test_visited_mass_depends_only_on_the_number_of_visited_outcomes():
    for n in range(1, 6):
        for visited in itertools.product((False, True), repeat=n):
            k = sum(visited)
            got = np.sqrt(_visited_mass(n, visited), dtype=np.float32)
            assert got.tobytes() == _table_entry(n, k).tobytes(), (n, visited)


def test_unvisited_terms_leave_the_sum_unchanged():
    # the lane-chain form used at a root that carries Dirichlet noise adds +0.0 for unvisited outcomes
    rng = np.random.default_rng(7)
    for _ in range(2000):
        n = int(rng.integers(1, 6))
        pri = rng.random(n).astype(np.float32)
        vis = rng.random(n) < 0.5
        a = np.float32(0.0)
        b = np.float32(0.0)
        for i in range(n):
            if vis[i]:
                a = np.float32(a + pri[i])
            b = np.float32(b + (pri[i] if vis[i] else np.float32(0.0)))
        assert a.tobytes() == b.tobytes()


def test_half_shared_memory_layout_fits_the_164_kb_carve_out():
    # half_smem_bytes(max_depth, batch_cap) of mcts_half.cuh for the bench configuration (50 turns -> depth 51, batch 16)
    def half_smem_bytes(max_depth: int, bc: int) -> int:
        b = 896 + bc * 8
        b += (max_depth * 4 + 15) & ~15
        b += bc * 32
        b += (bc + 1) * 8
        return (b + 15) & ~15

    per_half = half_smem_bytes(51, 16)
    assert per_half == 1888
    per_block = 2 * per_half + 176 + 1024  # one warp, static tables, per-block reserve
    assert 32 * per_block <= 164 * 1024  # 32 one-warp blocks per SM
    # the move table is five wide: a cell has at most four open directions plus STAY
    assert 64 * 5 * 2 == 640 and 640 + 64 * 4 == 896


def _pad_cnn_state_dict(sd: dict, cr: int, c: int = 64) -> dict:
    """The widening rule of nn_cnn.cu `pad_channels`, restated: every channel axis grows to `c` with zeros
    (BatchNorm running_var with ones), the combiner keeps its player columns after the 64 trunk columns."""
    out = {}
    for name, a in sd.items():
        a = np.asarray(a)
        if name == "stem.weight":
            b = np.zeros((c,) + a.shape[1:], a.dtype); b[:cr] = a
        elif name.endswith((".conv1.weight", ".conv2.weight")):
            b = np.zeros((c, c) + a.shape[2:], a.dtype); b[:cr, :cr] = a
        elif name.endswith(".pool_conv.weight"):
            b = np.zeros((a.shape[0], c) + a.shape[2:], a.dtype); b[:, :cr] = a
        elif name.endswith((".pool_linear.weight", ".pool_linear.bias")):
            b = np.zeros((c,) + a.shape[1:], a.dtype); b[:cr] = a
        elif name == "combiner.0.weight":
            b = np.zeros((a.shape[0], c + a.shape[1] - cr), a.dtype); b[:, :cr] = a[:, :cr]; b[:, c:] = a[:, cr:]
        elif a.ndim == 1 and a.shape[0] == cr and (name.startswith("stem_bn.") or ".bn1." in name or ".bn2." in name
                                                     or ".pool_bn." in name):
            b = np.full((c,), 1.0 if name.endswith(".running_var") else 0.0, a.dtype); b[:cr] = a
        else:
            b = a
        out[name] = b
    return out


def test_zero_padded_cnn_trunk_is_the_same_network():
    """The claim behind the CUDA loader's channel padding, checked with the fp32 restatement of PyRatCNN."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent))
    from nn_ref import cnn_forward, make_cnn_state_dict

    obs = np.load(Path(__file__).resolve().parent / "golden" / "flat_builder_7x7.npz")["obs"][:24]
    for cr, g in ((32, 16), (16, 16), (48, 32)):
        sd = make_cnn_state_dict(6, ("res", "gpool", "res"), channels=cr, gpool_channels=g)
        a = cnn_forward(sd, obs, 7, 7)
        b = cnn_forward(_pad_cnn_state_dict(sd, cr), obs, 7, 7)
        for x, y in zip(a, b):
            assert np.abs(x - y).max() <= 1e-6
