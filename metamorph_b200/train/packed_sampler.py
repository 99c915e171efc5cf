"""Token-budget batch sampler for packed training (second half).

The reference ships a `LengthGroupedSampler` (metamorph_trainer.py:105-154) that would sort mega-batches by length to
limit padding, but its scripts never enable it and every batch is padded to its longest sample
(metamorph_arch.py:361-399). With sequence packing (`model/interleave_plan.py::pack_plan`) the useful quantity is not
"similar lengths per batch" but "rows that are full": this sampler plans every step as `rows_per_batch` rows of
`pack_len` positions per rank and fills them with first-fit-decreasing bin packing over a shuffled window of samples.

  * deterministic: every rank derives the same global plan from (seed, epoch) — no communication;
  * each sample is used exactly once per epoch (the ragged tail is dropped or emitted as a short last step);
  * the rows of a step are dealt to the ranks so that their token counts are balanced (data-parallel ranks wait for
    the slowest one);
  * a rank's batch is emitted row by row, longest sample first, which is the order `pack_plan`'s next-fit walk turns
    back into (at most) the planned rows.
Lengths are INTERLEAVED lengths: text tokens with every image placeholder expanded to its visual tokens
(`interleaved_length`).
"""
from __future__ import annotations

from typing import Iterator, List, Sequence

import torch


def interleaved_length(n_tokens: int, n_images: int, image_tokens: int) -> int:
    """Positions a sample occupies after `prepare_inputs_labels_for_multimodal` (metamorph_arch.py:286-340): every
    image placeholder (one id) becomes `image_tokens` rows."""
    return n_tokens - n_images + n_images * image_tokens


def first_fit_decreasing(indices: Sequence[int], lengths: Sequence[int], capacity: int) -> List[List[int]]:
    """Rows (lists of sample indices, longest first) of total length <= capacity."""
    rows: List[List[int]] = []
    free: List[int] = []
    for i in sorted(indices, key=lambda j: (-lengths[j], j)):
        n = lengths[i]
        if n > capacity:
            raise ValueError(f"sample {i} has {n} positions, more than pack_len {capacity}")
        for r, f in enumerate(free):
            if n <= f:
                rows[r].append(i)
                free[r] -= n
                break
        else:
            rows.append([i])
            free.append(capacity - n)
    return rows


def next_fit_rows(lengths_in_order: Sequence[int], capacity: int) -> int:
    """Number of rows `pack_plan` (next-fit, order preserving) produces for samples of these lengths."""
    rows, used = 1, 0
    for n in lengths_in_order:
        if used + n > capacity and used > 0:
            rows += 1
            used = 0
        used += n
    return rows


class TokenBudgetBatchSampler(torch.utils.data.Sampler):
    """Yields, for this rank, the list of dataset indices of each step."""

    def __init__(self, lengths: Sequence[int], pack_len: int, rows_per_batch: int, world_size: int = 1, rank: int = 0,
                 seed: int = 0, window_rows: int = 16, drop_last: bool = True):
        if not 0 <= rank < world_size:
            raise ValueError("rank out of range")
        if any(n <= 0 for n in lengths):
            raise ValueError("lengths must be positive")
        self.lengths = list(lengths)
        self.pack_len, self.rows_per_batch = int(pack_len), int(rows_per_batch)
        self.world_size, self.rank, self.seed = world_size, rank, seed
        self.rows_per_step = self.rows_per_batch * world_size
        # bin packing looks at ~window_rows steps' worth of samples at a time: large enough to fill rows well, small
        # enough to keep the order random at the scale of the dataset
        mean = max(1.0, sum(self.lengths) / len(self.lengths))
        self.window = max(1, int(window_rows * self.rows_per_step * self.pack_len / mean))
        self.drop_last = drop_last
        self.epoch = 0
        self._plan_cache = None

    def set_epoch(self, epoch: int):
        self.epoch = epoch
        self._plan_cache = None

    # ------------------------------------------------------------------ global plan (identical on every rank)
    def _plan(self) -> List[List[List[int]]]:
        """steps -> rank -> flat index list (rows concatenated)."""
        if self._plan_cache is not None:
            return self._plan_cache
        g = torch.Generator().manual_seed(self.seed * 1_000_003 + self.epoch)
        order = torch.randperm(len(self.lengths), generator=g).tolist()
        rows: List[List[int]] = []
        for w0 in range(0, len(order), self.window):
            rows += first_fit_decreasing(order[w0:w0 + self.window], self.lengths, self.pack_len)
        # shuffle rows so that consecutive steps do not share a window's length profile
        perm = torch.randperm(len(rows), generator=g).tolist()
        rows = [rows[i] for i in perm]
        steps = []
        for s0 in range(0, len(rows), self.rows_per_step):
            chunk = rows[s0:s0 + self.rows_per_step]
            if len(chunk) < self.rows_per_step and self.drop_last:
                break
            # balance tokens across ranks: fullest rows first, each to the currently lightest rank with room
            fill = lambda row: sum(self.lengths[i] for i in row)  # noqa: E731
            per_rank: List[List[List[int]]] = [[] for _ in range(self.world_size)]
            load = [0] * self.world_size
            for row in sorted(chunk, key=lambda r: (-fill(r), r[0])):
                cands = [k for k in range(self.world_size) if len(per_rank[k]) < self.rows_per_batch]
                k = min(cands, key=lambda q: (load[q], q))
                per_rank[k].append(row)
                load[k] += fill(row)
            steps.append([[i for row in per_rank[k] for i in row] for k in range(self.world_size)])
        self._plan_cache = steps
        return steps

    def __iter__(self) -> Iterator[List[int]]:
        for step in self._plan():
            if step[self.rank]:
                yield step[self.rank]

    def __len__(self) -> int:
        return sum(1 for step in self._plan() if step[self.rank])

    # ------------------------------------------------------------------ diagnostics
    def efficiency(self) -> float:
        """Valid positions / positions processed (rows_per_batch * pack_len per rank and step) over the epoch."""
        steps = self._plan()
        valid = sum(self.lengths[i] for step in steps for idxs in step for i in idxs)
        return valid / max(1, len(steps) * self.rows_per_step * self.pack_len)
