"""Encoder halves of the reference autoencoders on the fused B200 kernel.

Mirrors the interface the retrieval path uses (retrieval/embedder.py:42-46, main.py:106-144
of the reference): an object with `.encode(x)`, `.eval()`, `.to(device)` and
`load_state_dict` of the reference checkpoints' key sets.  Decoders, `reparameterize`
and `forward` are training-only in the reference and are not provided.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_void_p
from typing import Dict, Mapping, Optional, Tuple

import numpy as np
import torch

from . import _native as nat

_KIND = {"dae": nat.LK_AE_DAE, "cae": nat.LK_AE_CAE, "vae": nat.LK_AE_VAE_MU}
_KEYS = {
    "dae": ("encoder.0.weight", "encoder.0.bias", "encoder.2.weight", "encoder.2.bias"),
    "cae": ("encoder.0.weight", "encoder.0.bias", "encoder.2.weight", "encoder.2.bias"),
    "vae": ("encoder.0.weight", "encoder.0.bias", "mu_layer.weight", "mu_layer.bias"),
}


class _FusedEncoder:
    kind = "dae"

    def __init__(self, input_dim: int, latent_dim: int, hidden_dim: int = 512, *, device: Optional[int] = None):
        self.input_dim, self.latent_dim, self.hidden_dim = int(input_dim), int(latent_dim), int(hidden_dim)
        self.device = device
        self._h = c_void_p()
        self._lib = None
        self._weights: Optional[Dict[str, np.ndarray]] = None
        self._kernel = "auto"
        self._precision = "fp32"
        self.training = False

    # -- nn.Module-shaped surface used by the reference pipeline -----------------------
    def eval(self):
        self.training = False
        return self

    def to(self, device):
        if device is not None and str(device).startswith("cuda"):
            idx = torch.device(device).index
            new = torch.cuda.current_device() if idx is None else idx
            if new != self.device:
                self.device = new
                self._release()
        return self

    def load_state_dict(self, state_dict: Mapping[str, object], strict: bool = True):
        w = {}
        shapes = ((self.hidden_dim, self.input_dim), (self.hidden_dim,), (self.latent_dim, self.hidden_dim),
                  (self.latent_dim,))
        for name, key, shape in zip(("w0", "b0", "w1", "b1"), _KEYS[self.kind], shapes):
            if key not in state_dict:
                raise KeyError(f"Missing key in state_dict: {key}")
            v = state_dict[key]
            v = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
            v = np.ascontiguousarray(v, dtype=np.float32)
            if tuple(v.shape) != shape:
                raise RuntimeError(f"size mismatch for {key}: checkpoint {tuple(v.shape)} vs model {shape}")
            w[name] = v
        self._weights = w
        self._release()
        return self

    def set_kernel(self, kernel: str):
        """"auto" (tensor cores for batches of >= 256 rows when the dims allow, fp32 FMA otherwise),
        "umma" (split-bf16 tcgen05 kernel: 3 MMAs per product, ~1e-5 of the row scale) or
        "simt" (fp32 FMA kernel, the reference's arithmetic)."""
        if kernel not in nat.KERNELS:
            raise ValueError(f"Unknown kernel: {kernel}")
        self._kernel = kernel
        if self._h:
            nat.check(self._lib.lk_ae_set_kernel(self._h, nat.KERNELS[kernel]), "lk_ae_set_kernel")
        return self

    def set_precision(self, precision: str):
        """"fp32" (default): fp32-level results on every kernel.  "bf16": inputs, weights and hidden
        activations rounded to bf16, fp32 accumulate, on the tensor-core kernel (about twice as fast;
        the search index stores the latents in bf16 anyway)."""
        if precision not in nat.STORAGE:
            raise ValueError(f"Unknown precision: {precision}")
        self._precision = precision
        if self._h:
            nat.check(self._lib.lk_ae_set_precision(self._h, nat.STORAGE[precision]), "lk_ae_set_precision")
        return self

    # -- native handle -------------------------------------------------------------------
    def _release(self):
        if self._h:
            self._lib.lk_ae_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self._release()
        except Exception:
            pass

    def _handle(self):
        if self._h:
            return self._h
        if self._weights is None:
            raise RuntimeError("load_state_dict must be called before encode")
        self._lib = nat.load()
        nat.require_device()
        if self.device is None:
            self.device = torch.cuda.current_device()
        w = self._weights
        nat.check(
            self._lib.lk_ae_create(byref(self._h), int(self.device), _KIND[self.kind], self.input_dim, self.hidden_dim,
                                   self.latent_dim, c_void_p(w["w0"].ctypes.data), c_void_p(w["b0"].ctypes.data),
                                   c_void_p(w["w1"].ctypes.data), c_void_p(w["b1"].ctypes.data)),
            "lk_ae_create",
        )
        if self._kernel != "auto":
            nat.check(self._lib.lk_ae_set_kernel(self._h, nat.KERNELS[self._kernel]), "lk_ae_set_kernel")
        if self._precision != "fp32":
            nat.check(self._lib.lk_ae_set_precision(self._h, nat.STORAGE[self._precision]), "lk_ae_set_precision")
        return self._h

    def _encode(self, x: torch.Tensor) -> torch.Tensor:
        """fp32 [M, input_dim] (CPU or CUDA) -> fp32 [M, latent_dim] on the same device."""
        h = self._handle()
        x = x.detach().to(torch.float32).contiguous()
        if x.dim() == 1:
            x = x.unsqueeze(0)
        if x.size(-1) != self.input_dim:
            raise ValueError(f"expected [..., {self.input_dim}] input, got {tuple(x.shape)}")
        lead = x.shape[:-1]
        x2 = x.reshape(-1, self.input_dim)
        if x2.is_cuda and x2.device.index != self.device:
            x2 = x2.to(f"cuda:{self.device}")
        # host rows -> host latents (retrieval/embedder.py:24-48 returns CPU fp32): page-locked, so that the
        # device -> host copies of the chunks run at link speed under the kernels of the next ones
        z = torch.empty((x2.size(0), self.latent_dim), dtype=torch.float32, device=x2.device,
                        pin_memory=not x2.is_cuda)
        mem = nat.LK_DEVICE if x2.is_cuda else nat.LK_HOST
        stream = int(torch.cuda.current_stream(self.device).cuda_stream)
        nat.check(self._lib.lk_ae_encode(h, c_void_p(x2.data_ptr()), mem, x2.size(0), c_void_p(z.data_ptr()), mem,
                                         c_void_p(stream)), "lk_ae_encode")
        return z.reshape(*lead, self.latent_dim)

    def encode(self, x: torch.Tensor):
        return self._encode(x)


class DenoisingAutoencoder(_FusedEncoder):
    """encode = Linear -> ReLU -> Linear (models/denoising_autoencoder.py:19-23,33-34)."""
    kind = "dae"


class ContrastiveAutoencoder(_FusedEncoder):
    """encode = Linear -> ReLU -> Linear -> L2 normalise (models/contrastive_autoencoder.py:10-14,23-25)."""
    kind = "cae"


class VariationalAutoencoder(_FusedEncoder):
    """encode -> (mu, logvar) in the reference (models/variational_autoencoder.py:26-30); retrieval keeps
    mu only (retrieval/embedder.py:44-45), so logvar is not computed and is returned as None."""
    kind = "vae"

    def encode(self, x: torch.Tensor) -> Tuple[torch.Tensor, None]:
        return self._encode(x), None


def load_autoencoder(kind: str, checkpoint, *, input_dim: int = 384, latent_dim: int = 64, hidden_dim: int = 512,
                     device: Optional[int] = None):
    """main.py:106-144 of the reference: build cls(input_dim, latent_dim, hidden_dim), load the
    checkpoint (a path to a .pth/.npz, or a state_dict mapping), eval()."""
    classes = {"vae": VariationalAutoencoder, "dae": DenoisingAutoencoder, "cae": ContrastiveAutoencoder,
               "contrastive": ContrastiveAutoencoder}
    if kind not in classes:
        raise ValueError(f"Unknown autoencoder type: {kind}")
    if isinstance(checkpoint, (str, bytes)) or hasattr(checkpoint, "__fspath__"):
        import os

        if not os.path.exists(checkpoint):
            raise FileNotFoundError(f"Checkpoint not found: {checkpoint}")  # main.py:139
        if str(checkpoint).endswith(".npz"):
            checkpoint = dict(np.load(checkpoint))
        else:
            checkpoint = torch.load(checkpoint, map_location="cpu")
    model = classes[kind](input_dim, latent_dim, hidden_dim, device=device)
    model.load_state_dict(checkpoint)
    return model.eval()
