"""Fused SimCLR info-NCE loss (SURVEY.md section 8 f, rank 4), backed by the sm_100a CUDA library.

Replaces, for the image-encoder pre-training of the reference (scripts/unimodel/
unimodel_training_for_image_encoder.py:26-58), the two lines of bioscanclip/util/simclr.py:118-119

    logits, labels = self.info_nce_loss(features)      # simclr.py:64-92: [M, M-1] logits, diagonal removed
    loss = self.criterion(logits, labels)              # nn.CrossEntropyLoss(), labels == 0

by ``loss = info_nce_loss(features, batch_size, n_views, temperature)``: same value and gradient, but the
M x M similarity matrix, its boolean masks and the re-ordered [M, M-1] logits are never materialised (same
tcgen05 / CUDA-core kernels as the contrastive loss with the main diagonal excluded; csrc/infonce.cu).

Like the reference trainer, the loss is local to the rank (SimCLR.train never gathers features).
There is no CPU or eager-PyTorch fallback: non-CUDA inputs raise.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _lib
from .loss import _DT, _select_path


class _InfoNCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, inv_temperature, path):
        lib = _lib.load()
        z = features.detach().contiguous()
        m, d = z.shape
        device, dtype = z.device, z.dtype
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        with torch.cuda.device(device):
            inv = torch.empty(m, dtype=torch.float32, device=device)
            _lib.check(lib.clibd_row_inv_norm(z.data_ptr(), _DT[dtype], m, d, inv.data_ptr(), stream))
            nbytes = lib.clibd_loss_scratch_bytes(m, m, d, path, 0)
            if nbytes < 0:
                raise ValueError("clibd_b200: bad info-NCE shape")
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=device)
            rowsum = torch.empty(m, dtype=torch.float32, device=device)
            loss = torch.empty((), dtype=torch.float32, device=device)
            _lib.check(lib.clibd_infonce_forward(z.data_ptr(), _DT[dtype], inv.data_ptr(), m, d, inv_temperature, path,
                                                 scratch.data_ptr(), nbytes, rowsum.data_ptr(), loss.data_ptr(),
                                                 stream))
        ctx.z, ctx.inv, ctx.scratch = z, inv, scratch
        ctx.meta = (m, d, inv_temperature, path, dtype, device)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        m, d, inv_temperature, path, dtype, device = ctx.meta
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        with torch.cuda.device(device):
            # grad_output (GradScaler's 65536, simclr.py:123) stays on the device
            g = grad_out.detach().to(device=device, dtype=torch.float32).reshape(1).contiguous()
            dz = torch.empty((m, d), dtype=dtype, device=device)
            _lib.check(lib.clibd_infonce_backward(ctx.z.data_ptr(), _DT[dtype], ctx.inv.data_ptr(), m, d,
                                                  inv_temperature, path, ctx.scratch.data_ptr(), ctx.scratch.numel(),
                                                  1.0, g.data_ptr(), dz.data_ptr(), stream))
        return dz, None, None


def info_nce_loss(features: torch.Tensor, batch_size: int, n_views: int = 2, temperature: float = 0.07,
                  tensor_core_operands=None) -> torch.Tensor:
    """``criterion(*SimCLR.info_nce_loss(features))`` (simclr.py:64-92, 119) as one fused loss.

    features: [n_views * batch_size, d]; rows v * batch_size + i are view v of image i (the trainer's
    ``torch.cat([images_1, images_2])``, simclr.py:111).  Returns the 0-d float32 mean cross-entropy.
    ``tensor_core_operands``: None (fp32 input -> exact CUDA-core path, bf16/fp16 input -> tcgen05 with that
    operand type), or "fp32" / "bf16" / "fp16" to force a path."""
    if n_views != 2:
        raise NotImplementedError("clibd_b200 fuses info-NCE for n_views == 2 (the only value the reference "
                                  "trainer can produce: it concatenates exactly two augmented batches)")
    if features.dim() != 2:
        raise ValueError("features must be a 2-D [n_views * batch_size, dim] tensor")
    if features.shape[0] != n_views * batch_size:
        # the reference fails here as well: its [M, M] label mask no longer matches the similarity matrix
        raise ValueError("features.shape[0] must equal n_views * batch_size")
    if batch_size < 1:
        raise ValueError("empty batch")
    if not features.is_cuda:
        raise RuntimeError("clibd_b200 runs on CUDA tensors only (there is no CPU fallback)")
    if not temperature > 0:
        raise ValueError("temperature must be positive")
    if features.dtype not in _DT:
        features = features.to(torch.float32)
    m, d = features.shape
    path = _select_path(features.dtype, tensor_core_operands, m if d <= 768 else 0, d)  # tcgen05 info-NCE: d <= 768
    return _InfoNCEFn.apply(features, 1.0 / float(temperature), path)


class InfoNCELoss(nn.Module):
    """Module form: ``InfoNCELoss(batch_size, n_views, temperature)(features)``."""

    def __init__(self, batch_size: int, n_views: int = 2, temperature: float = 0.07, tensor_core_operands=None):
        super().__init__()
        self.batch_size = batch_size
        self.n_views = n_views
        self.temperature = temperature
        self.tensor_core_operands = tensor_core_operands

    def forward(self, features):
        return info_nce_loss(features, self.batch_size, self.n_views, self.temperature, self.tensor_core_operands)
