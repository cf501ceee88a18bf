"""Game construction for the CUDA backend: `ar_game_pod` packing and a game generator.

Mirrors what the reference does inside Rust before self-play starts
(`make_games`, crates/alpharat-sampling/src/bindings.rs:489-533) and what `rust_mcts_search`
reads from a `PyRat` object (crates/alpharat-mcts/src/bindings.rs:250).  The third-party
engine's random generators are not available offline, so layouts come from this module's own
documented generator: open maze (or explicit walls/mud), corner starts, cheese
drawn without replacement with optional 180-degree rotational symmetry, PRNG = SplitMix64 keyed
by the game index.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any, Iterable, Sequence

import numpy as np

from ._native import AR_MAX_CELLS, GamePod

UP, RIGHT, DOWN, LEFT, STAY = 0, 1, 2, 3, 4
_DELTA = {UP: (0, 1), RIGHT: (1, 0), DOWN: (0, -1), LEFT: (-1, 0)}
_MASK64 = (1 << 64) - 1


class SplitMix64:
    """SplitMix64 stream (the generator's only source of randomness)."""

    def __init__(self, seed: int) -> None:
        self.state = seed & _MASK64

    def next(self) -> int:
        self.state = (self.state + 0x9E3779B97F4A7C15) & _MASK64
        z = self.state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK64
        return z ^ (z >> 31)

    def below(self, n: int) -> int:
        return self.next() % n

    def random(self) -> float:
        """Uniform in [0, 1) with 53 bits."""
        return (self.next() >> 11) * (1.0 / (1 << 53))


@dataclass
class GameSpec:
    """Explicit description of one PyRat position (host-side twin of `ar_game_pod`)."""

    width: int
    height: int
    max_turns: int
    p1: tuple[int, int]
    p2: tuple[int, int]
    cheese: Sequence[tuple[int, int]]
    walls: Sequence[tuple[tuple[int, int], tuple[int, int]]] = field(default_factory=list)
    mud: Sequence[tuple[tuple[int, int], tuple[int, int], int]] = field(default_factory=list)
    turn: int = 0
    p1_score: float = 0.0
    p2_score: float = 0.0
    p1_mud: int = 0
    p2_mud: int = 0


def _direction(a: tuple[int, int], b: tuple[int, int]) -> int:
    d = (b[0] - a[0], b[1] - a[1])
    for k, v in _DELTA.items():
        if v == d:
            return k
    raise ValueError(f"cells {a} and {b} are not adjacent")


def move_cost_table(width: int, height: int, walls: Iterable = (), mud: Iterable = ()) -> np.ndarray:
    """u8[cells*4]: 0 = wall/boundary, 1 = open, c >= 2 = mud cost (cell = y*width + x)."""
    t = np.ones((height, width, 4), dtype=np.uint8)
    t[:, 0, LEFT] = 0
    t[:, width - 1, RIGHT] = 0
    t[0, :, DOWN] = 0
    t[height - 1, :, UP] = 0
    for a, b in walls:
        d = _direction(tuple(a), tuple(b))
        t[a[1], a[0], d] = 0
        t[b[1], b[0], (d + 2) % 4] = 0
    for a, b, v in mud:
        d = _direction(tuple(a), tuple(b))
        t[a[1], a[0], d] = v
        t[b[1], b[0], (d + 2) % 4] = v
    return t.reshape(-1)


def pack_pod(spec: GameSpec, out: GamePod | None = None) -> GamePod:
    cells = spec.width * spec.height
    if cells > AR_MAX_CELLS:
        raise ValueError(f"{spec.width}x{spec.height} exceeds AR_MAX_CELLS={AR_MAX_CELLS}")
    pod = out if out is not None else GamePod()
    C.memset(C.byref(pod), 0, C.sizeof(pod))
    pod.width, pod.height = spec.width, spec.height
    pod.p1_x, pod.p1_y = spec.p1
    pod.p2_x, pod.p2_y = spec.p2
    pod.p1_mud, pod.p2_mud = spec.p1_mud, spec.p2_mud
    pod.turn, pod.max_turns = spec.turn, spec.max_turns
    pod.p1_score, pod.p2_score = spec.p1_score, spec.p2_score
    mc = move_cost_table(spec.width, spec.height, spec.walls, spec.mud)
    C.memmove(pod.move_cost, mc.ctypes.data, mc.size)
    for x, y in spec.cheese:
        c = y * spec.width + x
        pod.cheese[c >> 3] |= 1 << (c & 7)
    return pod


def pods_array(specs: Sequence[GameSpec]):
    arr = (GamePod * len(specs))()
    for i, s in enumerate(specs):
        pack_pod(s, arr[i])
    return arr


def pod_from_pyrat(game: Any) -> GamePod:
    """Pack a duck-typed `PyRat` (CLAUDE.md "PyRat Game API") into an `ar_game_pod`."""

    def xy(p: Any) -> tuple[int, int]:
        return (int(p.x), int(p.y)) if hasattr(p, "x") else (int(p[0]), int(p[1]))

    spec = GameSpec(
        width=int(game.width),
        height=int(game.height),
        max_turns=int(game.max_turns),
        p1=xy(game.player1_position),
        p2=xy(game.player2_position),
        cheese=[xy(c) for c in game.cheese_positions()],
        walls=[(xy(w.pos1), xy(w.pos2)) for w in game.wall_entries()],
        mud=[(xy(m.pos1), xy(m.pos2), int(m.value)) for m in game.mud_entries()],
        turn=int(game.turn),
        p1_score=float(game.player1_score),
        p2_score=float(game.player2_score),
        p1_mud=int(game.player1_mud_turns),
        p2_mud=int(game.player2_mud_turns),
    )
    return pack_pod(spec)


def random_cheese(width: int, height: int, count: int, symmetric: bool, rng: SplitMix64,
                  exclude: Iterable[tuple[int, int]]) -> list[tuple[int, int]]:
    """`count` distinct cells, none in `exclude`; 180-degree symmetric pairs when asked."""
    excl = set(exclude)
    chosen: list[tuple[int, int]] = []
    taken = set(excl)
    centre = ((width - 1) / 2, (height - 1) / 2)
    has_centre = width % 2 == 1 and height % 2 == 1
    centre_cell = (width // 2, height // 2)
    if symmetric:
        if count % 2 == 1:
            if not has_centre or centre_cell in taken:
                raise ValueError("odd symmetric cheese count needs a free centre cell")
            chosen.append(centre_cell)
            taken.add(centre_cell)
        while len(chosen) < count:
            c = rng.below(width * height)
            p = (c % width, c // width)
            q = (width - 1 - p[0], height - 1 - p[1])
            if p in taken or q in taken or p == q or (p[0], p[1]) == centre:
                continue
            chosen += [p, q]
            taken.update((p, q))
    else:
        while len(chosen) < count:
            c = rng.below(width * height)
            p = (c % width, c // width)
            if p in taken:
                continue
            chosen.append(p)
            taken.add(p)
    return chosen


def random_maze(width: int, height: int, wall_density: float, mud_density: float, mud_range: int,
                symmetric: bool, rng: SplitMix64):
    """Connected random maze: (walls, mud) as `GameSpec` edge lists.

    Same knobs as the engine's `MazeParams` (bindings.rs:509-519: wall_density, mud_density,
    mud_range, connected=True, symmetric).  The engine's own generator is third-party code that is
    not available offline, so this is this module's documented procedure, not a bit-for-bit twin:
      1. every inner edge (orbit of the 180-degree rotation when `symmetric`) is walled with
         probability `wall_density`, visited in row-major order (horizontal edge before vertical);
      2. while the maze is disconnected, the walled orbits that separate two components are listed
         in the same order and one of them, drawn uniformly, is opened;
      3. every open orbit gets mud with probability `mud_density`, cost uniform in [2, mud_range].
    """
    def mirror(c):
        return (width - 1 - c[0], height - 1 - c[1])

    orbits, seen = [], set()
    for y in range(height):
        for x in range(width):
            for nx, ny in ((x + 1, y), (x, y + 1)):
                if nx >= width or ny >= height:
                    continue
                e = ((x, y), (nx, ny))
                if e in seen:
                    continue
                orbit = [e]
                if symmetric:
                    m = (mirror(e[1]), mirror(e[0]))  # keeps the (lower, upper) orientation
                    if m != e:
                        orbit.append(m)
                seen.update(orbit)
                orbits.append(orbit)
    walled = [rng.random() < wall_density for _ in orbits]

    def components():
        parent = list(range(width * height))

        def find(a):
            while parent[a] != a:
                parent[a] = parent[parent[a]]
                a = parent[a]
            return a

        for o, w in zip(orbits, walled):
            if not w:
                for a, b in o:
                    ra, rb = find(a[1] * width + a[0]), find(b[1] * width + b[0])
                    if ra != rb:
                        parent[ra] = rb
        return find

    while True:
        find = components()
        roots = {find(c) for c in range(width * height)}
        if len(roots) == 1:
            break
        cand = [i for i, (o, w) in enumerate(zip(orbits, walled)) if w and any(
            find(a[1] * width + a[0]) != find(b[1] * width + b[0]) for a, b in o)]
        walled[cand[rng.below(len(cand))]] = False
    walls, mud = [], []
    for o, w in zip(orbits, walled):
        if w:
            walls.extend(o)
        elif mud_density > 0.0 and rng.random() < mud_density:
            cost = 2 + (rng.below(mud_range - 1) if mud_range > 2 else 0)
            mud.extend((a, b, cost) for a, b in o)
    return walls, mud


def make_games(
    num_games: int,
    *,
    width: int,
    height: int,
    cheese_count: int,
    max_turns: int,
    cheese_symmetric: bool = True,
    maze_type: str = "open",
    positions: str = "corners",
    wall_density: float = 0.7,
    mud_density: float = 0.1,
    maze_symmetric: bool = True,
    first_index: int = 0,
    layout_seed: int | None = None,
) -> list[GameSpec]:
    """Generator behind `cuda_self_play` (same axes as `make_games`, bindings.rs:489-533):
    maze_type open / classic (wall 0.7, mud 0.1, symmetric) / random; positions corners / random
    (P2 is the 180-degree mirror of P1 when the cheese is symmetric); random cheese.
    One SplitMix64 stream per game.  `layout_seed=None` keys the stream by the game index alone
    (reproducible: tests, benchmarks); a run seed is mixed in otherwise, so that successive sampling
    runs see different boards, as the reference's `config.create(None)` does (bindings.rs:529-532)."""
    if maze_type not in ("open", "classic", "random"):
        raise ValueError(f"unknown maze_type: {maze_type!r}")
    if positions not in ("corners", "random"):
        raise ValueError(f"unknown positions: {positions!r}")
    out = []
    for i in range(num_games):
        key = first_index + i
        if layout_seed is not None:  # one SplitMix64 output of the run seed, xor-folded with the index
            key = (SplitMix64(layout_seed & ((1 << 64) - 1)).next() ^ (key * 0x9E3779B97F4A7C15)) & ((1 << 64) - 1)
        rng = SplitMix64(key)
        walls, mud = [], []
        if maze_type == "classic":
            walls, mud = random_maze(width, height, 0.7, 0.1, 3, True, rng)
        elif maze_type == "random":
            walls, mud = random_maze(width, height, wall_density, mud_density, 3 if mud_density > 0.0 else 2,
                                     maze_symmetric, rng)
        if positions == "corners":
            p1, p2 = (0, 0), (width - 1, height - 1)
        else:
            while True:
                c = rng.below(width * height)
                p1 = (c % width, c // width)
                if cheese_symmetric:
                    p2 = (width - 1 - p1[0], height - 1 - p1[1])
                else:
                    c2 = rng.below(width * height)
                    p2 = (c2 % width, c2 // width)
                if p1 != p2:
                    break
        cheese = random_cheese(width, height, cheese_count, cheese_symmetric, rng, (p1, p2))
        out.append(GameSpec(width, height, max_turns, p1, p2, cheese, walls=walls, mud=mud))
    return out


def maze_array(spec: GameSpec) -> np.ndarray:
    """i8[H, W, 4]: -1 wall, 1 open, >= 2 mud (build_maze_array, selfplay.rs:374-392)."""
    mc = move_cost_table(spec.width, spec.height, spec.walls, spec.mud).astype(np.int16)
    mc[mc == 0] = -1
    return mc.astype(np.int8).reshape(spec.height, spec.width, 4)
