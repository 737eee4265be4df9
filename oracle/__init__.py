"""CPU oracle for the SearchTransfer hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain numpy (and, for the timed CPU baseline, plain torch-CPU)
restatement of the reference algorithm in /root/reference/model/SearchTransfer.py
and of the three fusion lines of /root/reference/model/speinet.py::_decode, plus (SURVEY.md
section 8(f) row 3) the Richardson-Lucy edge prior of /root/reference/model/rcl.py:22-51.

Nothing in the product package `speinet_b200` imports it.  The only allowed
importers are `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py`, and there only as the checker or the
timed CPU baseline -- never as the thing shipped.

Parity pinning: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the reference module itself, executed
in the build container by `tests/golden/make_golden.py` and committed as small
fixtures under `tests/golden/*.npz`; `tests/test_oracle.py` re-checks the oracle
against those fixtures on every run.
"""
from .search_transfer_np import (  # noqa: F401
    unfold, fold, l2_normalize, relevance, search_transfer, self_transfer_S,
    closed_form_transfer, near_tie_agreement,
)
from .fusion_np import bicubic_upsample, conv1x1, fuse_level  # noqa: F401
from .rl_deconv_np import create_blur_kernel, r_l_per_channel  # noqa: F401
