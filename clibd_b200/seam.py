"""The two seams either side of the hot path (SURVEY.md section 8 f, ranks 1-3), device-resident.

  * ``softmax_mean`` / ``SoftmaxMeanHead``  -- ``logits.softmax(dim=-1).mean(dim=1)``, the BarcodeBERT head of
    bioscanclip/model/dna_encoder.py:137, as ONE forward and ONE backward kernel (the reference runs a
    softmax that writes [n, 133, 768] probabilities, a mean that reads them back, and the matching
    multi-pass autograd).
  * ``EmbeddingStore`` / ``get_feature_and_label``  -- bioscanclip/epoch/inference_epoch.py:42-125: the
    per-batch ``F.normalize(...).cpu().tolist()`` + final ``np.array`` round trip becomes one normalise-and-
    append kernel into a float32 device store.
  * ``derived_feature_types`` / ``get_features_and_label``  -- bioscanclip/util/util.py:702-742: averaged,
    concatenated and all-key feature construction on the device (float64 where numpy computes in float64,
    so the retrieval that follows sees the reference's values).

The dictionaries these functions return feed ``clibd_b200.inference_and_print_result`` directly
(device tensors instead of numpy arrays).  No CPU fallback: non-CUDA tensors raise.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _lib

_DT = {torch.float32: _lib.DT_F32, torch.bfloat16: _lib.DT_BF16, torch.float16: _lib.DT_F16}


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"clibd_b200: {what} runs on CUDA tensors only (there is no CPU fallback)")


def _aligned_contiguous(t: torch.Tensor) -> torch.Tensor:
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


# ------------------------------------------------------------------------------------------------
# BarcodeBERT head
# ------------------------------------------------------------------------------------------------
class _SoftmaxMeanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits):
        lib = _lib.load()
        x = _aligned_contiguous(logits.detach())
        n, t, c = x.shape
        out = torch.empty((n, c), dtype=x.dtype, device=x.device)
        nbytes = lib.clibd_softmax_mean_scratch_bytes(n, t, c, _DT[x.dtype])
        if nbytes < 0:
            raise ValueError("clibd_b200: bad softmax_mean shape")
        scratch = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.clibd_softmax_mean_forward(x.data_ptr(), _DT[x.dtype], n, t, c, out.data_ptr(),
                                                      scratch.data_ptr(), int(nbytes), _stream(x.device)))
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        (x,) = ctx.saved_tensors
        n, t, c = x.shape
        g = _aligned_contiguous(grad_out.detach().to(x.dtype))
        dx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(lib.clibd_softmax_mean_backward(x.data_ptr(), g.data_ptr(), _DT[x.dtype], n, t, c,
                                                       dx.data_ptr(), _stream(x.device)))
        return dx


def softmax_mean(logits: torch.Tensor) -> torch.Tensor:
    """``logits.softmax(dim=-1).mean(dim=1)`` for [n, tokens, classes] logits (dna_encoder.py:137)."""
    if logits.dim() != 3:
        raise ValueError("softmax_mean expects [batch, tokens, classes] logits")
    _require_cuda(logits, "softmax_mean")
    if logits.dtype not in _DT:
        raise TypeError(f"softmax_mean supports float32 / bfloat16 / float16 logits, not {logits.dtype}")
    if logits.shape[0] == 0:
        return logits.new_zeros((0, logits.shape[2]))
    return _SoftmaxMeanFn.apply(logits)


class SoftmaxMeanHead(nn.Module):
    """Drop-in for the tail of ``CLIBDDNAEncoder.forward`` (dna_encoder.py:131-137):
    ``SoftmaxMeanHead()(self.base_dna_encoder(sequence).logits)``."""

    def forward(self, logits: torch.Tensor) -> torch.Tensor:
        return softmax_mean(logits)


# ------------------------------------------------------------------------------------------------
# embedding hand-off
# ------------------------------------------------------------------------------------------------
class EmbeddingStore:
    """Growing float32 [rows, d] device store of L2-normalised embeddings (one kernel per appended batch)."""

    def __init__(self, capacity: int = 0, device=None):
        self.device = torch.device(device) if device is not None else None
        self.capacity = int(capacity)
        self.buf = None
        self.rows = 0

    def _reserve(self, need: int, d: int, device):
        if self.buf is None:
            cap = max(self.capacity, need, 1024)
            self.buf = torch.empty((cap, d), dtype=torch.float32, device=device)
            self.device = device
        elif need > self.buf.shape[0]:
            cap = max(need, 2 * self.buf.shape[0])
            grown = torch.empty((cap, self.buf.shape[1]), dtype=torch.float32, device=self.buf.device)
            grown[: self.rows] = self.buf[: self.rows]
            self.buf = grown

    def append(self, features: torch.Tensor):
        """F.normalize(features, dim=-1) appended as float32 rows (inference_epoch.py:96-101)."""
        if features.dim() != 2:
            raise ValueError("features must be [batch, dim]")
        _require_cuda(features, "EmbeddingStore.append")
        x = features.detach()
        if x.dtype not in _DT:
            x = x.float()
        x = x.contiguous()
        n, d = x.shape
        if self.buf is not None and d != self.buf.shape[1]:
            raise ValueError("feature width changed between batches")
        self._reserve(self.rows + n, d, x.device)
        lib = _lib.load()
        with torch.cuda.device(x.device):
            _lib.check(lib.clibd_embed_append(x.data_ptr(), _DT[x.dtype], n, d, self.buf.data_ptr(),
                                              self.buf.shape[0], self.buf.stride(0), self.rows, _stream(x.device)))
        self.rows += n

    def tensor(self):
        """[rows, d] float32 view of what has been appended (None if nothing was)."""
        return None if self.buf is None or self.rows == 0 else self.buf[: self.rows]


def convert_label_dict_to_list_of_dict(label_batch):
    """inference_epoch.py:8-20."""
    return [{"order": o, "family": f, "genus": g, "species": s}
            for o, f, g, s in zip(label_batch["order"], label_batch["family"], label_batch["genus"],
                                  label_batch["species"])]


def get_feature_and_label(dataloader, model, device, for_open_clip=False, multi_gpu=False, dna_tokenizer=None):
    """inference_epoch.py:42-125 with device-resident outputs.

    Returns (file_name_list, image_features, dna_features, text_features, label_list); each feature entry is a
    float32 [n, d] DEVICE tensor (the reference returns float64 numpy arrays holding the same float32
    values) or None.  String DNA batches are tokenised with ``dna_tokenizer`` (the reference loads
    "bioscan-ml/BarcodeBERT" from the hub, inference_epoch.py:53; pass the tokenizer in -- there is no network
    access here) using the reference's arguments (padding='max_length', truncation=True, max_length=133)."""
    stores = [EmbeddingStore(device=device), EmbeddingStore(device=device), EmbeddingStore(device=device)]
    label_list, file_name_list = [], []
    model.eval()
    with torch.no_grad():
        for batch in dataloader:
            processid_batch, image_input_batch, dna_input_batch, input_ids, token_type_ids, attention_mask, label_batch = batch
            if for_open_clip:
                language_input = input_ids
            else:
                language_input = {"input_ids": input_ids.to(device), "token_type_ids": token_type_ids.to(device),
                                  "attention_mask": attention_mask.to(device)}
            if isinstance(dna_input_batch, torch.Tensor):
                dna_input_batch = dna_input_batch.to(device)
            else:
                if dna_tokenizer is None:
                    raise ValueError("string DNA batches need dna_tokenizer= (the BarcodeBERT tokenizer)")
                toks = [dna_tokenizer(seq, padding="max_length", truncation=True, max_length=133,
                                      return_tensors="pt")["input_ids"] for seq in dna_input_batch]
                dna_input_batch = torch.stack(toks).squeeze(1).to(device)
            image_output, dna_output, language_output, _logit_scale, _logit_bias = model(
                image_input_batch.to(device), dna_input_batch, language_input)
            for store, out in zip(stores, (image_output, dna_output, language_output)):
                if out is not None:
                    store.append(out)
            label_list.extend(convert_label_dict_to_list_of_dict(label_batch))
            file_name_list.extend(list(processid_batch))
    return file_name_list, stores[0].tensor(), stores[1].tensor(), stores[2].tensor(), label_list


def derived_feature_types(image, dna, text=None, label_list=None, for_key_set=False):
    """util.py:711-737 on device tensors: ``averaged_feature`` = np.mean([image, dna], axis=0) (float64, like numpy
    on the reference's float64 arrays), ``concatenated_feature`` = [image | dna], and for key sets
    ``all_key_features`` = vstack(image, dna, text) with the label list repeated three times."""
    out = {"averaged_feature": None, "concatenated_feature": None, "all_key_features": None,
           "all_key_features_label": None}
    if image is not None and dna is not None:
        _require_cuda(image, "derived_feature_types")
        out["averaged_feature"] = (image.double() + dna.double()) / 2.0
        out["concatenated_feature"] = torch.cat((image, dna), dim=1)
    if for_key_set and image is not None and dna is not None and text is not None:
        out["all_key_features"] = torch.cat((image, dna, text), dim=0)
        out["all_key_features_label"] = list(label_list) * 3 if label_list is not None else None
    return out


def get_features_and_label(dataloader, model, device, for_key_set=False, for_open_clip=False, dna_tokenizer=None):
    """util.py:702-742: the dictionary of one split, every feature entry a device tensor."""
    model.eval()
    file_name_list, image, dna, text, label_list = get_feature_and_label(
        dataloader, model, device, for_open_clip=for_open_clip, multi_gpu=False, dna_tokenizer=dna_tokenizer)
    derived = derived_feature_types(image, dna, text, label_list, for_key_set=for_key_set)
    return {
        "file_name_list": file_name_list,
        "encoded_dna_feature": dna,
        "encoded_image_feature": image,
        "encoded_language_feature": text,
        "averaged_feature": derived["averaged_feature"],
        "concatenated_feature": derived["concatenated_feature"],
        "label_list": label_list,
        "all_key_features": derived["all_key_features"],
        "all_key_features_label": derived["all_key_features_label"],
    }
