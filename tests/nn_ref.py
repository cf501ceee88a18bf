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


# ---------------------------------------------------------------------------------------------
# SymmetricMLP (alpharat/nn/models/symmetric.py:124-229) and PyRatCNN (alpharat/nn/models/cnn/
# model.py:117-230, blocks.py, heads.py) restated in numpy, with seeded state_dicts under the
# reference's own parameter names.
# ---------------------------------------------------------------------------------------------
def _bn_params(r, sd, name, ch):
    f = np.float32
    sd[f"{name}.weight"] = (1.0 + 0.2 * r.standard_normal(ch)).astype(f)
    sd[f"{name}.bias"] = (0.1 * r.standard_normal(ch)).astype(f)
    sd[f"{name}.running_mean"] = (0.3 * r.standard_normal(ch)).astype(f)
    sd[f"{name}.running_var"] = (0.5 + r.random(ch)).astype(f)
    sd[f"{name}.num_batches_tracked"] = np.array(7, dtype=np.int64)


def make_symmetric_state_dict(seed: int, width: int, height: int, hidden: int = 256) -> dict[str, np.ndarray]:
    r = np.random.default_rng(seed)
    f = np.float32
    S = width * height

    def lin(o, i, scale):
        return (r.standard_normal((o, i)) * scale).astype(f), (r.standard_normal(o) * 0.1).astype(f)

    sd = {}
    sd["shared_encoder.0.weight"], sd["shared_encoder.0.bias"] = lin(hidden, 5 * S + 1, (2.0 / (5 * S + 1)) ** 0.5)
    _bn_params(r, sd, "shared_encoder.1", hidden)
    sd["player_encoder.0.weight"], sd["player_encoder.0.bias"] = lin(hidden, S + 2, (2.0 / (S + 2)) ** 0.5)
    _bn_params(r, sd, "player_encoder.1", hidden)
    sd["trunk.0.weight"], sd["trunk.0.bias"] = lin(hidden, 2 * hidden, (1.0 / hidden) ** 0.5)
    _bn_params(r, sd, "trunk.1", hidden)
    sd["trunk.4.weight"], sd["trunk.4.bias"] = lin(hidden, hidden, (2.0 / hidden) ** 0.5)
    _bn_params(r, sd, "trunk.5", hidden)
    sd["policy_head.weight"], sd["policy_head.bias"] = lin(5, 2 * hidden, 0.15)
    sd["value_head.weight"], sd["value_head.bias"] = lin(1, 2 * hidden, 0.15)
    return sd


def _softmax(z):
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)


def _softplus(v):
    return np.where(v > 20, v, np.log1p(np.exp(np.minimum(v, 20)))).astype(np.float32)


def symmetric_forward(sd, obs: np.ndarray, width: int, height: int):
    S = width * height
    x = obs.astype(np.float32)
    sc = 7 * S
    maze, p1pos, p2pos, cheese = x[:, :4 * S], x[:, 4 * S:5 * S], x[:, 5 * S:6 * S], x[:, 6 * S:7 * S]
    shared_raw = np.concatenate([maze, cheese, x[:, sc + 1:sc + 2]], axis=1)
    p1_raw = np.concatenate([p1pos, x[:, sc + 2:sc + 3], x[:, sc + 4:sc + 5]], axis=1)
    p2_raw = np.concatenate([p2pos, x[:, sc + 3:sc + 4], x[:, sc + 5:sc + 6]], axis=1)
    ws, bs = _fold(sd, "shared_encoder.0", "shared_encoder.1")
    wp, bp = _fold(sd, "player_encoder.0", "player_encoder.1")
    w1, b1 = _fold(sd, "trunk.0", "trunk.1")
    w2, b2 = _fold(sd, "trunk.4", "trunk.5")
    relu = lambda a: np.maximum(a, 0).astype(np.float32)
    shared = relu(shared_raw @ ws.T + bs)
    hs = []
    for raw in (p1_raw, p2_raw):
        p = relu(raw @ wp.T + bp)
        t = relu(np.concatenate([shared, p], axis=1) @ w1.T + b1)
        hs.append(relu(t @ w2.T + b2))
    agg = hs[0] + hs[1]
    outs = []
    for h in hs:
        c = np.concatenate([h, agg], axis=1)
        outs.append((_softmax(c @ sd["policy_head.weight"].T + sd["policy_head.bias"]),
                     _softplus(c @ sd["value_head.weight"].T + sd["value_head.bias"])[:, 0]))
    return outs[0][0], outs[1][0], outs[0][1], outs[1][1]


def make_cnn_state_dict(seed: int, blocks=("res", "res", "gpool"), channels: int = 64, gpool_channels: int = 32,
                        player_dim: int = 32, hidden_dim: int = 64) -> dict[str, np.ndarray]:
    r = np.random.default_rng(seed)
    f = np.float32
    C = channels

    def conv(o, i, k):
        return (r.standard_normal((o, i, k, k)) * (2.0 / (i * k * k)) ** 0.5).astype(f)

    def lin(o, i, scale):
        return (r.standard_normal((o, i)) * scale).astype(f), (r.standard_normal(o) * 0.1).astype(f)

    sd = {"stem.weight": conv(C, 5, 3)}
    _bn_params(r, sd, "stem_bn", C)
    for b, kind in enumerate(blocks):
        pre = f"blocks.{b}"
        _bn_params(r, sd, f"{pre}.bn1", C)
        sd[f"{pre}.conv1.weight"] = conv(C, C, 3)
        _bn_params(r, sd, f"{pre}.bn2", C)
        sd[f"{pre}.conv2.weight"] = (conv(C, C, 3) * 0.5).astype(f)
        if kind == "gpool":
            _bn_params(r, sd, f"{pre}.pool_bn", C)
            sd[f"{pre}.pool_conv.weight"] = conv(gpool_channels, C, 1)
            sd[f"{pre}.pool_linear.weight"], sd[f"{pre}.pool_linear.bias"] = lin(C, 2 * gpool_channels, 0.1)
    sd["player_encoder.0.weight"], sd["player_encoder.0.bias"] = lin(player_dim, 3, 0.8)
    sd["combiner.0.weight"], sd["combiner.0.bias"] = lin(hidden_dim, C + player_dim, (2.0 / (C + player_dim)) ** 0.5)
    sd["policy_head.linear.weight"], sd["policy_head.linear.bias"] = lin(5, 2 * hidden_dim, 0.2)
    sd["value_head.linear.weight"], sd["value_head.linear.bias"] = lin(1, 2 * hidden_dim, 0.2)
    return sd


def _bn2d(sd, name, x):
    s = (sd[f"{name}.weight"] / np.sqrt(sd[f"{name}.running_var"] + np.float32(1e-5))).astype(np.float32)
    t = (sd[f"{name}.bias"] - sd[f"{name}.running_mean"] * s).astype(np.float32)
    return x * s[None, :, None, None] + t[None, :, None, None]


def _conv3x3(x, w):
    """x [B, Ci, H, W], w [Co, Ci, 3, 3], padding 1 (cross-correlation, as torch)."""
    B, Ci, H, W = x.shape
    xp = np.zeros((B, Ci, H + 2, W + 2), np.float32)
    xp[:, :, 1:-1, 1:-1] = x
    out = np.zeros((B, w.shape[0], H, W), np.float32)
    for ky in range(3):
        for kx in range(3):
            out += np.einsum("bchw,oc->bohw", xp[:, :, ky:ky + H, kx:kx + W], w[:, :, ky, kx], optimize=True)
    return out.astype(np.float32)


def cnn_forward(sd, obs: np.ndarray, width: int, height: int):
    S = width * height
    x = obs.astype(np.float32)
    B = x.shape[0]
    sc = 7 * S
    maze = x[:, :4 * S].reshape(B, height, width, 4).transpose(0, 3, 1, 2)
    cheese = x[:, 6 * S:7 * S].reshape(B, 1, height, width)
    spatial = np.concatenate([maze, cheese], axis=1)
    p1_mask, p2_mask = x[:, 4 * S:5 * S], x[:, 5 * S:6 * S]
    prog = x[:, sc + 1:sc + 2]
    sides = [np.concatenate([x[:, sc + 4:sc + 5], x[:, sc + 2:sc + 3], prog], axis=1),
             np.concatenate([x[:, sc + 5:sc + 6], x[:, sc + 3:sc + 4], prog], axis=1)]
    relu = lambda a: np.maximum(a, 0).astype(np.float32)
    f = relu(_bn2d(sd, "stem_bn", _conv3x3(spatial, sd["stem.weight"])))
    b = 0
    while f"blocks.{b}.conv1.weight" in sd:
        pre = f"blocks.{b}"
        reg = _conv3x3(relu(_bn2d(sd, f"{pre}.bn1", f)), sd[f"{pre}.conv1.weight"])
        reg = _conv3x3(relu(_bn2d(sd, f"{pre}.bn2", reg)), sd[f"{pre}.conv2.weight"])
        out = reg + f
        if f"{pre}.pool_conv.weight" in sd:
            pool = relu(_bn2d(sd, f"{pre}.pool_bn", f))
            pool = np.einsum("bchw,oc->bohw", pool, sd[f"{pre}.pool_conv.weight"][:, :, 0, 0], optimize=True)
            cat = np.concatenate([pool.mean(axis=(2, 3)), pool.max(axis=(2, 3))], axis=1).astype(np.float32)
            po = cat @ sd[f"{pre}.pool_linear.weight"].T + sd[f"{pre}.pool_linear.bias"]
            out = out + po[:, :, None, None]
        f = out.astype(np.float32)
        b += 1
    ff = f.reshape(B, f.shape[1], -1)
    hs = []
    for mask, side in zip((p1_mask, p2_mask), sides):
        feat = (ff * mask[:, None, :]).sum(axis=2)
        e = relu(side @ sd["player_encoder.0.weight"].T + sd["player_encoder.0.bias"])
        hs.append(relu(np.concatenate([feat, e], axis=1) @ sd["combiner.0.weight"].T + sd["combiner.0.bias"]))
    agg = hs[0] + hs[1]
    outs = []
    for h in hs:
        c = np.concatenate([h, agg], axis=1)
        outs.append((_softmax(c @ sd["policy_head.linear.weight"].T + sd["policy_head.linear.bias"]),
                     _softplus(c @ sd["value_head.linear.weight"].T + sd["value_head.linear.bias"])[:, 0]))
    return outs[0][0], outs[1][0], outs[0][1], outs[1][1]
