"""CLAP text tower on the device: drop-in for the ``clap_encoder`` argument of ``AudioTextHTDemucs``
(/root/reference/src/models/stem_separation/ATHTDemucs_v2.py:151-161, used at :238-248).

``ClapTextModelWithProjectionB200`` mirrors ``transformers.ClapTextModelWithProjection`` for this path:
``forward(input_ids=..., attention_mask=...).text_embeds`` ([P, 512], not normalised); ``ClapModelTextB200`` adds
``get_text_features(**inputs)`` (L2-normalised, the ``ClapModel`` branch of ``_get_clap_embeddings`` with transformers 4.51).
Weights are loaded from the HF state_dict keys ``text_model.*`` / ``text_projection.*`` (audio-tower keys are ignored).
The tokenizer stays the caller's (a vocabulary file, not arithmetic).  fp32, no CPU fallback.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch

from . import lib as _lib


class ClapTextModelWithProjectionB200(torch.nn.Module):
    def __init__(self):
        super().__init__()
        lib = _lib.load()
        self._table = [(lib.athtd_clap_param_name(i).decode(), lib.athtd_clap_param_numel(i), lib.athtd_clap_param_offset(i))
                       for i in range(lib.athtd_clap_param_count())]
        self._total = lib.athtd_clap_params_total()
        self.register_buffer("flat_params", torch.zeros(self._total, dtype=torch.float32), persistent=False)
        self._loaded = False

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        """HF keys ``{prefix}text_model.*`` / ``{prefix}text_projection.*`` -> the flat fp32 buffer (works for a direct
        ``tower.load_state_dict(hf_sd)`` and for a parent checkpoint whose keys start with ``clap.``); audio-tower keys
        of a full ``ClapModel`` checkpoint are ignored."""
        found = 0
        for name, numel, off in self._table:
            t = state_dict.get(prefix + name)
            if t is None:
                missing_keys.append(prefix + name)
                continue
            if t.numel() != numel:
                error_msgs.append(f"{prefix + name}: expected {numel} elements, got {tuple(t.shape)}")
                continue
            self.flat_params[off:off + numel].copy_(t.detach().reshape(-1).float())
            found += 1
        if found == 0:
            del missing_keys[-len(self._table):]      # a checkpoint without CLAP weights: keep what is loaded, report nothing
        self._loaded = self._loaded or found == len(self._table)

    @torch.no_grad()
    def _embed(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], normalize: bool) -> torch.Tensor:
        if not self.flat_params.is_cuda:
            raise _lib.AthtdError("the CLAP text tower runs on CUDA only: call .to('cuda') (no CPU fallback)")
        dev = self.flat_params.device
        ids = input_ids.to(dev, torch.int64).contiguous()
        if ids.dim() != 2:
            raise ValueError("input_ids must be [P, S]")
        mask = torch.ones_like(ids) if attention_mask is None else attention_mask.to(dev, torch.int64).contiguous()
        P, S = ids.shape
        lib = _lib.load()
        ws = torch.empty(lib.athtd_clap_workspace_bytes(P, S), dtype=torch.uint8, device=dev)
        out = torch.empty(P, 512, dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.athtd_clap_text_forward(self.flat_params.data_ptr(), ids.data_ptr(), mask.data_ptr(), P, S, ws.data_ptr(),
                                               out.data_ptr(), 1 if normalize else 0, st), "athtd_clap_text_forward")
        return out

    def forward(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None, **_unused):
        return SimpleNamespace(text_embeds=self._embed(input_ids, attention_mask, False))


class ClapModelTextB200(ClapTextModelWithProjectionB200):
    """The text half of ``transformers.ClapModel`` as this path uses it (``get_text_features``)."""

    def get_text_features(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None, **_unused) -> torch.Tensor:
        return self._embed(input_ids, attention_mask, True)
