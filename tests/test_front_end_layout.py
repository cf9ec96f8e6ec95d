"""Host-side layout of the per-frame parameter row (cvae.py:196-202, global_optimization.py:96-115, :454): the fused
front-end node must cut the 78-D row exactly where body_params_encapsulate_batch cuts the 75-D row."""
import importlib

import torch

prior = importlib.import_module("4dcapture-fpv_b200.prior")


def test_row78_blocks_match_the_75d_column_split():
    row75 = torch.arange(75, dtype=torch.float32).repeat(2, 1)
    parts = prior.body_params_encapsulate_batch(row75)               # pure slicing: runs on the CPU
    assert tuple(parts) == prior._ROW78_KEYS
    widths75 = [parts[k].shape[1] for k in prior._ROW78_KEYS]
    assert sum(prior._ROW78) == 78 and sum(widths75) == 75
    # identical widths except the orientation block, which is 6-D in the optimised row and 3-D after decoding
    assert [w if k != "global_orient" else 3 for k, w in zip(prior._ROW78_KEYS, prior._ROW78)] == widths75
    assert prior._ROW78[prior._ROW78_KEYS.index("global_orient")] == 6
    # the blocks are contiguous and in column order
    start = 0
    for k in prior._ROW78_KEYS:
        assert float(parts[k][0, 0]) == start
        start += parts[k].shape[1]


def test_front_end_split_has_no_cpu_fallback():
    import pytest
    with pytest.raises(RuntimeError):
        prior.front_end_split(torch.zeros(4, 78))
