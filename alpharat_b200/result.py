"""Canonical search result — same dataclass as the reference (alpharat/mcts/result.py:16-43)."""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class SearchResult:
    """All arrays are in 5-action space [UP, RIGHT, DOWN, LEFT, STAY]; blocked actions are 0."""

    policy_p1: np.ndarray
    policy_p2: np.ndarray
    value_p1: float
    value_p2: float
    visit_counts_p1: np.ndarray
    visit_counts_p2: np.ndarray
    prior_p1: np.ndarray
    prior_p2: np.ndarray
    total_visits: int
