"""CPU restatement of the reference autoencoders' *encoder* forwards.  TEST INFRASTRUCTURE.

Follows (paths relative to /root/reference):
  models/denoising_autoencoder.py:19-23,33-34     z = W1 relu(W0 x + b0) + b1
  models/contrastive_autoencoder.py:10-14,23-25   the same, then F.normalize(z, dim=-1)
  models/variational_autoencoder.py:11-16,26-30   h = relu(W0 x + b0); mu = Wmu h + bmu
  retrieval/embedder.py:42-46                     retrieval keeps mu only (tuple -> [0])
"""
from __future__ import annotations

from typing import Dict, Mapping

import numpy as np
import torch
import torch.nn.functional as F

KINDS = ("dae", "cae", "vae")

# state_dict keys of the encoder half, per kind (decoder.* / logvar_layer.* are unused
# at retrieval time)
_KEYS = {
    "dae": ("encoder.0.weight", "encoder.0.bias", "encoder.2.weight", "encoder.2.bias"),
    "cae": ("encoder.0.weight", "encoder.0.bias", "encoder.2.weight", "encoder.2.bias"),
    "vae": ("encoder.0.weight", "encoder.0.bias", "mu_layer.weight", "mu_layer.bias"),
}


def load_encoder_weights(state_dict: Mapping[str, object], kind: str) -> Dict[str, torch.Tensor]:
    """Pick (w0 [H,D], b0 [H], w1 [Z,H], b1 [Z]) out of a reference state_dict
    (or out of the npz fixture written by tests/golden/make_golden.py)."""
    if kind not in _KEYS:
        raise ValueError(f"Unknown autoencoder kind: {kind}")
    names = ("w0", "b0", "w1", "b1")
    out = {}
    for name, key in zip(names, _KEYS[kind]):
        v = state_dict[key]
        out[name] = torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v).to(torch.float32).cpu()
    return out


def ae_encode(x: torch.Tensor, w: Mapping[str, torch.Tensor], kind: str, precision: str = "fp32") -> torch.Tensor:
    """Latent code used for retrieval: fp32 [M, Z].  For the VAE this is `mu`.
    precision="bf16": the reference's arithmetic fed bf16-rounded inputs, weights and hidden
    activations (what the engine's opt-in bf16 encoder computes; fp32 accumulation, fp32 biases)."""
    x = x.detach().to("cpu", torch.float32)
    r = (lambda t: t.to(torch.bfloat16).to(torch.float32)) if precision == "bf16" else (lambda t: t)
    h = torch.relu(F.linear(r(x), r(w["w0"]), w["b0"]))
    z = F.linear(r(h), r(w["w1"]), w["b1"])
    if kind == "cae":
        z = F.normalize(z, p=2, dim=-1)
    return z
