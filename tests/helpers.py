"""Shared helpers for the parity tests."""
import json
import os

import numpy as np
import torch

from list_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    """Returns (inputs regenerated from the recipe, npz dict of reference outputs)."""
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    kw = json.loads(str(z["recipe"]))
    inp = synth.make_inputs(**kw)
    # the regenerated inputs must be the ones the reference saw
    assert abs(inp.points.double().sum().item() - float(z["points_sum"])) < 1e-9 if "points_sum" in z else True
    return inp, z, kw


def singular_mask(inp, rel=1e-4):
    """Points whose perspective divide is ill-conditioned (SURVEY.md §7 'hard parts'):
    |h2 + 1e-8| tiny relative to the terms that formed it, so a 1-ulp difference in the 4x3
    product flips the clamp side.  Such points are reported, not compared."""
    q = (inp.points[:, :, [2, 1, 0]] * 2).double()
    T = inp.trans_mat.double()
    terms = torch.stack([q[..., 0] * T[:, None, 0, 2].squeeze(1)[:, None] if False else q[..., 0] * T[:, 0, 2][:, None],
                         q[..., 1] * T[:, 1, 2][:, None],
                         q[..., 2] * T[:, 2, 2][:, None],
                         T[:, 3, 2][:, None].expand_as(q[..., 0])], dim=-1)
    h2 = terms.sum(-1) + 1e-8
    return h2.abs() < rel * terms.abs().sum(-1)
