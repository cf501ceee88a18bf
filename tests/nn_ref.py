"""Plain fp32 restatement of the leaf evaluators' forward pass (test infrastructure).

`mlp_forward` follows PyRatMLP.predict (alpharat/nn/models/mlp.py:120-153) in eval mode; with
`emulate_bf16=True` it rounds inputs, folded weights and hidden activations to bf16 exactly where
the CUDA kernel does, which isolates layout bugs from precision effects.
Weights come from a seeded numpy generator so that the golden script (which feeds them to the real
reference class) and the GPU tests build identical models without shipping the weights.
"""

from __future__ import annotations

import numpy as np


def make_mlp_state_dict(seed: int, obs_dim: int, hidden: int = 256) -> dict[str, np.ndarray]:
    r = np.random.default_rng(seed)
    f = np.float32

    def lin(o, i, scale):
        return (r.standard_normal((o, i)) * scale).astype(f), (r.standard_normal(o) * 0.1).astype(f)

    sd = {}
    sd["trunk.0.weight"], sd["trunk.0.bias"] = lin(hidden, obs_dim, (2.0 / obs_dim) ** 0.5)
    sd["trunk.4.weight"], sd["trunk.4.bias"] = lin(hidden, hidden, (2.0 / hidden) ** 0.5)
    for bn in ("trunk.1", "trunk.5"):  # non-trivial BatchNorm statistics to exercise the folding
        sd[f"{bn}.weight"] = (1.0 + 0.2 * r.standard_normal(hidden)).astype(f)
        sd[f"{bn}.bias"] = (0.1 * r.standard_normal(hidden)).astype(f)
        sd[f"{bn}.running_mean"] = (0.3 * r.standard_normal(hidden)).astype(f)
        sd[f"{bn}.running_var"] = (0.5 + r.random(hidden)).astype(f)
        sd[f"{bn}.num_batches_tracked"] = np.array(7, dtype=np.int64)
    sd["policy_p1_head.weight"], sd["policy_p1_head.bias"] = lin(5, hidden, 0.2)
    sd["policy_p2_head.weight"], sd["policy_p2_head.bias"] = lin(5, hidden, 0.2)
    sd["value_head.weight"], sd["value_head.bias"] = lin(2, hidden, 0.2)
    return sd


def to_bf16(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even f32 -> bf16 -> f32."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    lsb = (u >> 16) & 1
    u = (u + 0x7FFF + lsb) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(np.shape(x))


def _fold(sd, lin, bn):
    w, b = sd[f"{lin}.weight"].astype(np.float32), sd[f"{lin}.bias"].astype(np.float32)
    s = (sd[f"{bn}.weight"] / np.sqrt(sd[f"{bn}.running_var"] + np.float32(1e-5))).astype(np.float32)
    return (w * s[:, None]).astype(np.float32), ((b - sd[f"{bn}.running_mean"]) * s + sd[f"{bn}.bias"]).astype(np.float32)


def mlp_forward(sd: dict[str, np.ndarray], obs: np.ndarray, emulate_bf16: bool = False):
    q = to_bf16 if emulate_bf16 else (lambda a: a)
    w1, b1 = _fold(sd, "trunk.0", "trunk.1")
    w2, b2 = _fold(sd, "trunk.4", "trunk.5")
    x = q(obs.astype(np.float32))
    h = np.maximum(x @ q(w1).T + b1, 0).astype(np.float32)
    h = np.maximum(q(h) @ q(w2).T + b2, 0).astype(np.float32)
    h = q(h)

    def softmax(z):
        z = z - z.max(axis=1, keepdims=True)
        e = np.exp(z)
        return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)

    p1 = softmax(h @ q(sd["policy_p1_head.weight"]).T + sd["policy_p1_head.bias"])
    p2 = softmax(h @ q(sd["policy_p2_head.weight"]).T + sd["policy_p2_head.bias"])
    v = h @ q(sd["value_head.weight"]).T + sd["value_head.bias"]
    v = np.where(v > 20, v, np.log1p(np.exp(np.minimum(v, 20)))).astype(np.float32)
    return p1, p2, v[:, 0], v[:, 1]


def random_positions(n: int, width: int, height: int, seed: int):
    """Mid-game positions with walls/mud/scores/mud timers for evaluator tests."""
    from alpharat_b200.games import GameSpec

    r = np.random.default_rng(seed)
    specs = []
    for _ in range(n):
        cells = [(x, y) for y in range(height) for x in range(width)]
        k = int(r.integers(1, 11))
        cheese = [cells[i] for i in r.choice(len(cells), size=k, replace=False)]
        walls, mud = [], []
        for _ in range(int(r.integers(0, 4))):
            x, y = int(r.integers(0, width - 1)), int(r.integers(0, height))
            walls.append(((x, y), (x + 1, y)))
        for _ in range(int(r.integers(0, 3))):
            x, y = int(r.integers(0, width)), int(r.integers(0, height - 1))
            if ((x, y), (x, y + 1)) not in walls:
                mud.append(((x, y), (x, y + 1), int(r.integers(2, 6))))
        max_turns = int(r.integers(20, 101))
        specs.append(GameSpec(
            width, height, max_turns,
            (int(r.integers(0, width)), int(r.integers(0, height))),
            (int(r.integers(0, width)), int(r.integers(0, height))),
            cheese, walls=walls, mud=mud, turn=int(r.integers(0, max_turns)),
            p1_score=float(r.integers(0, 9)) / 2, p2_score=float(r.integers(0, 9)) / 2,
            p1_mud=int(r.integers(0, 4)) if r.random() < 0.3 else 0,
            p2_mud=int(r.integers(0, 4)) if r.random() < 0.3 else 0))
    return specs
