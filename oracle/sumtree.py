"""NumPy restatement of the reference's float64 sum tree (TEST INFRASTRUCTURE).

Follows ``priority_tree.py:4-45`` of the reference:
  * layout      ``priority_tree.py:5-13``  array-backed complete binary tree,
                ``num_layers`` = smallest L with 2**(L-1) >= capacity, leaves
                start at node 2**(L-1)-1.
  * update      ``priority_tree.py:15-24`` leaf <- td**alpha (computed in the
                dtype of ``td``: float32 in the learner), then every ancestor
                level is RECOMPUTED from its two children (never delta-added).
                Duplicate indices: the last write wins (NumPy fancy assignment).
  * sample      ``priority_tree.py:26-45`` stratified prefix-sum descent in
                float64 and IS weights (p / min p)**-beta over the drawn batch.

The one liberty taken: ``sample`` can be handed the unit uniforms ``r`` in
[0,1) instead of drawing them, so that a GPU implementation can be compared
bit-for-bit on identical draws.  ``np.random.uniform(0, w, n)`` of the legacy
generator is exactly ``w * random_sample(n)``, which is what is computed here.
"""
from __future__ import annotations

import numpy as np


def num_layers_for(capacity: int) -> int:
    """priority_tree.py:6-8."""
    layers = 1
    while capacity > (1 << (layers - 1)):
        layers += 1
    return layers


class SumTreeOracle:
    def __init__(self, capacity: int, prio_exponent: float, is_exponent: float):
        self.capacity = int(capacity)
        self.num_layers = num_layers_for(capacity)
        self.leaf_base = (1 << (self.num_layers - 1)) - 1
        self.ptree = np.zeros((1 << self.num_layers) - 1, dtype=np.float64)
        self.prio_exponent = prio_exponent
        self.is_exponent = is_exponent

    # -- priority_tree.py:15-24 -------------------------------------------------
    def update(self, idxes: np.ndarray, td_error: np.ndarray) -> None:
        leaf_values = np.asarray(td_error) ** self.prio_exponent  # f32 stays f32
        self.set_leaves(idxes, leaf_values)

    def set_leaves(self, idxes: np.ndarray, leaf_values: np.ndarray) -> None:
        """Write already-exponentiated leaf priorities and rebuild ancestors."""
        nodes = np.asarray(idxes, dtype=np.int64) + self.leaf_base
        self.ptree[nodes] = leaf_values          # duplicates: last one wins
        for _ in range(self.num_layers - 1):
            nodes = np.unique((nodes - 1) // 2)
            self.ptree[nodes] = self.ptree[2 * nodes + 1] + self.ptree[2 * nodes + 2]

    # -- priority_tree.py:26-45 -------------------------------------------------
    def sample(self, num_samples: int, unit_uniforms: np.ndarray | None = None):
        total = self.ptree[0]
        interval = total / num_samples
        if unit_uniforms is None:
            unit_uniforms = np.random.random_sample(num_samples)
        unit_uniforms = np.asarray(unit_uniforms, dtype=np.float64)
        # np.arange(0, total, interval)[i] == i*interval; uniform(0, w) == w*r
        prefix = np.arange(num_samples, dtype=np.float64) * interval + interval * unit_uniforms

        node = np.zeros(num_samples, dtype=np.int64)
        for _ in range(self.num_layers - 1):
            left = self.ptree[2 * node + 1]
            go_left = prefix < left
            prefix = np.where(go_left, prefix, prefix - left)
            node = np.where(go_left, 2 * node + 1, 2 * node + 2)

        prio = self.ptree[node]
        is_weights = np.power(prio / prio.min(), -self.is_exponent)
        return node - self.leaf_base, is_weights

    @property
    def leaves(self) -> np.ndarray:
        return self.ptree[self.leaf_base:]
