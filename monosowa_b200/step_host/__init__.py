"""Host-side stalls of the MonoDETR training step, moved to the device.

The reference step spends ~50 ms of a 205 ms step inside ``SetCriterion`` with the GPU mostly idle
(profiles/r01b_train_breakdown.json): three Hungarian matchings that copy the cost matrix to the host and call
scipy 176 times each (matcher.py:87-104), two Python loops over every ground-truth box with four implicit
device synchronisations per box (ddn_loss.py:56-62, balancer.py:76-79), and a hand-written AdamW that issues
nine tiny kernels per parameter (optimizer_helper.py:76-127).  ``install(criterion, optimizer)`` swaps those
host sections for device-resident equivalents with the same results; nothing in the reference tree is edited
and the MSDA operator (monosowa_b200.ops) is independent of this module.

    from monosowa_b200 import step_host
    step_host.install(criterion, optimizer)        # after build_monodetr(...) / build_optimizer(...)
"""
from __future__ import annotations

from .lsa import group_lsa  # noqa: F401
from .patches import (DeviceMatcher, foreach_adamw_step, fuse_dense_attention, install, paint_depth_targets,  # noqa: F401
                      paint_foreground, pairwise_l1)
from .frozen_bn import frozen_bn_act, fuse_frozen_bn  # noqa: F401
