"""Prioritized sampling of initial states for ``mc_pilco(..., prioritized_replay=True)``.

Host-side counterpart of the reference's ``utils.SumTree`` (reference utils/experience_dataset.py:271-367)
with the same public behaviour -- ``append / update / renormalize / sample`` and the attributes the
algorithm reads (``size, max_size, max_p, counts``) -- so that seeded runs draw the same initial
states as the reference.  The priorities come from the fused reverse sweep: the norm of the TOTAL
dL/da_t of every imagined step (``pmb_rollout_backward(..., da_total)``), which is what the reference
collects with hooks on ``actions[t]`` (algorithms/mc_pilco.py:160-188).

Layout: a complete binary tree in one array, ``capacity - 1`` inner nodes followed by ``capacity``
leaves; node ``i`` has children ``2i + 1`` and ``2i + 2``; an inner node holds the sum of its children.
"""
import numpy as np


class SumTree:
    def __init__(self, max_size):
        self.max_size = int(max_size)
        self.data = [None] * self.max_size
        self.sum_tree = np.zeros(2 * self.max_size - 1)
        self.counts = np.zeros(self.max_size)
        self.idx = 0                # next leaf slot (ring buffer)
        self.size = 0
        self.max_p = 1.0            # largest priority ever assigned: new states enter with it
        self.max_count = 0
        self.norm_factor = 1.0      # priorities are stored scaled so that the root stays 1 after renormalize()

    # ---- writes -------------------------------------------------------------------------------
    def _leaf(self, slot):
        return slot + self.max_size - 1

    def append(self, item, priority):
        slot = self.idx
        self.data[slot] = item
        self.counts[slot] = 1
        self.update(self._leaf(slot), priority)
        self.idx = (slot + 1) % self.max_size
        self.size = min(self.size + 1, self.max_size)

    def update(self, node, priority):
        """Set the (unnormalised) priority of tree node ``node`` (a leaf index) and refresh its ancestors."""
        self.sum_tree[node] = priority * self.norm_factor
        while node:
            node = (node - 1) // 2
            self.sum_tree[node] = self.sum_tree[2 * node + 1] + self.sum_tree[2 * node + 2]
        self.max_p = max(self.max_p, priority)

    def renormalize(self):
        scale = 1.0 / self.sum_tree[0]
        self.norm_factor *= scale
        self.sum_tree *= scale

    # ---- reads --------------------------------------------------------------------------------
    def _descend(self, mass):
        """Leaf whose cumulative-priority interval contains ``mass`` (scalar walk)."""
        node, last = 0, len(self.sum_tree)
        while 2 * node + 1 < last:
            left = 2 * node + 1
            if mass <= self.sum_tree[left]:
                node = left
            else:
                mass -= self.sum_tree[left]
                node = left + 1
        return node

    def _descend_many(self, mass):
        """Vectorised walk for a whole batch of masses."""
        mass = np.array(mass, dtype=np.float64)
        node = np.zeros(len(mass), dtype=np.int64)
        last = len(self.sum_tree)
        left = 2 * node + 1
        live = left < last
        while live.any():
            lv = self.sum_tree[left]
            go_left = mass <= lv
            node = np.where(go_left, left, left + 1)
            mass = np.where(go_left, mass, mass - lv)
            left = 2 * node + 1
            live = left < last
            left = np.where(live, left, node)
        return node

    def get(self, mass):
        node = self._descend(mass)
        return [node, self.sum_tree[node], self.data[node - self.max_size + 1]]

    def get_batch(self, mass):
        nodes = self._descend_many(np.atleast_1d(mass))
        return nodes, self.sum_tree[nodes], [self.data[i] for i in nodes - self.max_size + 1]

    def sample(self, batchsize, beta=1.0):
        """Stratified sample: one draw per equal slice of the total priority mass.  Returns
        (items, tree node indices, importance weights normalised to max 1)."""
        total = self.sum_tree[0]
        mass = (np.arange(batchsize) + np.random.rand(batchsize)) * (total / batchsize)
        if batchsize < 32:
            picked = [self.get(m) for m in mass]
            nodes = np.array([q[0] for q in picked])
            prio = [q[1] for q in picked]
            items = [q[2] for q in picked]
        else:
            nodes, prio, items = self.get_batch(mass)
        slots = nodes - self.max_size + 1
        self.counts[slots] += 1
        self.max_count = max(self.max_count, self.counts[slots].max())
        weights = (self.size * (np.array(prio) / total)) ** -beta
        return items, nodes, weights / weights.max()
