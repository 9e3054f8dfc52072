"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import numpy as np
import torch

from depgan_b200 import synth
from oracle import depgan_oracle as O


def gen_weights(nicg=1, nc_out=1, seed=0, trained_like=True):
    return synth.init_weights(O.gen_manifest(nicg, nc_out), seed=seed, trained_like=trained_like)


def critic_weights(h, w, seed=1, trained_like=True):
    return synth.init_weights(O.critic_manifest(h, w), seed=seed, trained_like=trained_like)


def oracle_gen(P, x, z, head="tanh", dtype=torch.float64):
    Pt = O.to_torch(P, dtype)
    with torch.no_grad():
        return O.gen_forward(Pt, torch.as_tensor(x, dtype=dtype), torch.as_tensor(z, dtype=dtype), head).numpy()


def oracle_critic(P, x, dtype=torch.float64):
    Pt = O.to_torch(P, dtype)
    with torch.no_grad():
        return O.critic_forward(Pt, torch.as_tensor(x, dtype=dtype)).numpy()
