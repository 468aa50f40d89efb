"""Host-side mirror of the reference's module API for the hot path.

``AudioTextHTDemucsB200`` keeps the constructor signature, sub-module names, parameter shapes and
therefore the ``state_dict`` layout of ``AudioTextHTDemucs``
(/root/reference/src/models/stem_separation/ATHTDemucs_v2.py:142-188, printed tree
AudioTextHTDemucs_Full.txt:3-629), so ``load_state_dict(ckpt["model_state_dict"], strict=False)``
(benchmark.py:144-145) works unchanged.  The sub-modules below are PARAMETER CONTAINERS: none of
them implements a PyTorch forward.  ``forward`` hands device pointers to libathtd.so (hand-written
sm_100a kernels); there is no eager / CPU fallback.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Union

import torch
import torch.nn as nn

from .engine import Engine
from .lib import AthtdError

ENC_CH = [48, 96, 192, 384]
DEC_CH = [384, 192, 96, 48, 4]


class _NoForward(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise AthtdError(f"{type(self).__name__} only holds parameters; compute runs in libathtd.so")


class LayerScaleParams(_NoForward):
    def __init__(self, channels: int, init: float):
        super().__init__()
        self.scale = nn.Parameter(torch.full((channels,), float(init)))


class DConvParams(_NoForward):
    """state_dict names of demucs DConv: layers.{d}.{0,1,3,4}.{weight,bias}, layers.{d}.6.scale."""

    def __init__(self, channels: int, compress: int = 8, depth: int = 2, init: float = 1e-3):
        super().__init__()
        hidden = channels // compress
        self.layers = nn.ModuleList()
        for d in range(depth):
            self.layers.append(nn.Sequential(
                nn.Conv1d(channels, hidden, 3, dilation=2 ** d, padding=2 ** d), nn.GroupNorm(1, hidden), nn.Identity(),
                nn.Conv1d(hidden, 2 * channels, 1), nn.GroupNorm(1, 2 * channels), nn.Identity(),
                LayerScaleParams(channels, init)))


class HEncLayerParams(_NoForward):
    def __init__(self, chin: int, chout: int, freq: bool):
        super().__init__()
        self.freq = freq
        self.empty = False
        if freq:
            self.conv = nn.Conv2d(chin, chout, (8, 1), (4, 1), (2, 0))
            self.rewrite = nn.Conv2d(chout, 2 * chout, 1)
        else:
            self.conv = nn.Conv1d(chin, chout, 8, 4, 2)
            self.rewrite = nn.Conv1d(chout, 2 * chout, 1)
        self.dconv = DConvParams(chout)


class _Embedding(_NoForward):
    def __init__(self):
        super().__init__()
        self.embedding = nn.Embedding(512, 48)


class _AttnParams(_NoForward):
    """nn.MultiheadAttention key names: in_proj_weight, in_proj_bias, out_proj.{weight,bias}."""

    def __init__(self, dim: int):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * dim, dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * dim))
        self.out_proj = nn.Linear(dim, dim)
        nn.init.xavier_uniform_(self.in_proj_weight)


class _XfLayerParams(_NoForward):
    def __init__(self, cross: bool, dim: int = 512, hidden: int = 2048, init: float = 1e-4):
        super().__init__()
        setattr(self, "cross_attn" if cross else "self_attn", _AttnParams(dim))
        self.linear1 = nn.Linear(dim, hidden)
        self.linear2 = nn.Linear(hidden, dim)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        if cross:
            self.norm3 = nn.LayerNorm(dim)
        self.norm_out = nn.GroupNorm(1, dim)
        self.gamma_1 = LayerScaleParams(dim, init)
        self.gamma_2 = LayerScaleParams(dim, init)


class CrossTransformerParams(_NoForward):
    def __init__(self):
        super().__init__()
        self.norm_in = nn.LayerNorm(512)
        self.norm_in_t = nn.LayerNorm(512)
        self.layers = nn.ModuleList([_XfLayerParams(i % 2 == 1) for i in range(5)])
        self.layers_t = nn.ModuleList([_XfLayerParams(i % 2 == 1) for i in range(5)])


class HTDemucsParams(_NoForward):
    """Parameter tree with the key names of demucs.htdemucs.HTDemucs that the hot path reads
    (ATHTDemucs_v2.py:197-234).  Pass a real demucs ``HTDemucs`` instead when demucs is installed:
    only its parameters are read, its forward is never called."""

    def __init__(self):
        super().__init__()
        self.encoder = nn.ModuleList()
        self.tencoder = nn.ModuleList()
        for i in range(4):
            self.encoder.append(HEncLayerParams(4 if i == 0 else ENC_CH[i - 1], ENC_CH[i], True))
            self.tencoder.append(HEncLayerParams(2 if i == 0 else ENC_CH[i - 1], ENC_CH[i], False))
        self.freq_emb = _Embedding()
        self.freq_emb_scale = 0.2
        self.channel_upsampler = nn.Conv1d(384, 512, 1)
        self.channel_downsampler = nn.Conv1d(512, 384, 1)
        self.channel_upsampler_t = nn.Conv1d(384, 512, 1)
        self.channel_downsampler_t = nn.Conv1d(512, 384, 1)
        self.crosstransformer = CrossTransformerParams()
        self.bottom_channels = 512


class TextCrossAttention(_NoForward):
    """Parameters of ATHTDemucs_v2.py:21-36."""

    def __init__(self, feat_dim: int, text_dim: int, n_heads: int = 8):
        super().__init__()
        self.q_proj = nn.Linear(feat_dim, feat_dim)
        self.k_proj = nn.Linear(text_dim, feat_dim)
        self.v_proj = nn.Linear(text_dim, feat_dim)
        self.attn = _AttnParams(feat_dim)
        self.out_mlp = nn.Sequential(nn.Linear(feat_dim, feat_dim), nn.Identity(), nn.Linear(feat_dim, feat_dim))
        self.norm_q = nn.LayerNorm(feat_dim)
        self.norm_out = nn.LayerNorm(feat_dim)


class _DecoderParams(_NoForward):
    def __init__(self, channels: Sequence[int], two_d: bool):
        super().__init__()
        self.layers = nn.ModuleList()
        for i in range(len(channels) - 1):
            last = i == len(channels) - 2
            ct = (nn.ConvTranspose2d(channels[i], channels[i + 1], (8, 1), (4, 1), (2, 0)) if two_d
                  else nn.ConvTranspose1d(channels[i], channels[i + 1], 8, 4, 2))
            self.layers.append(nn.Sequential(ct, nn.Identity() if last else nn.GroupNorm(1, channels[i + 1]), nn.Identity()))


class FreqDecoder(_DecoderParams):
    """Parameters of ATHTDemucs_v2.py:61-80."""

    def __init__(self, channels: Sequence[int]):
        super().__init__(channels, True)


class TimeDecoder(_DecoderParams):
    """Parameters of ATHTDemucs_v2.py:107-123."""

    def __init__(self, channels: Sequence[int]):
        super().__init__(channels, False)


class AudioTextHTDemucsB200(nn.Module):
    """Drop-in for ``AudioTextHTDemucs`` (ATHTDemucs_v2.py:142-326) on one B200.

    forward(wav[B,2,T] float32 cuda, text) -> [B,2,T]; ``text`` is a list of B prompt strings, a
    bare string (B == 1, as test_inference.py:120 does) or a (B,512) tensor of text embeddings.
    Extra, B200-first entry point: ``separate_batch(wav, emb[B,P,512]) -> [B,P,2,T]`` encodes each
    segment once and decodes it for P prompts.
    """

    def __init__(self, htdemucs_model: Optional[nn.Module] = None, clap_encoder: Optional[nn.Module] = None,
                 clap_tokenizer=None, model_dim: int = 384, text_dim: int = 512, num_heads: int = 8,
                 sample_rate: int = 44100, segment: float = 7.8, precision: str = "bf16"):
        super().__init__()
        if model_dim != 384 or text_dim != 512 or num_heads != 8:
            raise ValueError("the B200 path implements the reference configuration model_dim=384, text_dim=512, heads=8")
        self.htdemucs = htdemucs_model if htdemucs_model is not None else HTDemucsParams()
        self.clap = clap_encoder if clap_encoder is not None else nn.Module()
        self.tokenizer = clap_tokenizer
        self.sample_rate = sample_rate
        self.segment = segment
        self.precision = precision
        for p in self.htdemucs.parameters():
            p.requires_grad = False
        for p in self.clap.parameters():
            p.requires_grad = False
        self.text_attn = TextCrossAttention(model_dim, text_dim, num_heads)
        self.freq_decoder = FreqDecoder(DEC_CH)
        self.time_decoder = TimeDecoder(DEC_CH)
        self.freq_out = nn.Conv2d(4, 2, 1)
        self.time_out = nn.Conv1d(4, 2, 1)
        self._engine: Optional[Engine] = None
        self._sig = None
        self._prompt_cache: Dict[str, torch.Tensor] = {}

    # ------------------------------------------------------------------ parameters -> device blob
    def _live_named(self) -> Dict[str, torch.Tensor]:
        live = {}
        for k, v in self.named_parameters():
            if not (k.startswith("clap.") or k.startswith("htdemucs.decoder.") or k.startswith("htdemucs.tdecoder.")):
                live[k] = v
        return live

    def engine(self, device=None) -> Engine:
        dev = torch.device(device) if device is not None else next(self.text_attn.parameters()).device
        if dev.type != "cuda":
            raise AthtdError("AudioTextHTDemucsB200 runs on a CUDA sm_100a device only (no CPU fallback)")
        if self._engine is None or self._engine.device != dev or self._engine.dtype != self.precision:
            self._engine = Engine(dev, self.precision)
            self._sig = None
        live = self._live_named()
        sig = tuple((v.data_ptr(), v._version) for v in live.values())
        if sig != self._sig:
            self._engine.load_params(live)
            self._sig = sig
        return self._engine

    def refresh(self) -> None:
        """Force re-upload / re-pack of the weights (after in-place edits that keep _version)."""
        self._sig = None

    # ------------------------------------------------------------------ text
    def register_prompt_embedding(self, prompt: str, emb: torch.Tensor) -> None:
        self._prompt_cache[prompt] = emb.detach().reshape(512).float()

    def _get_clap_embeddings(self, text: List[str], device) -> torch.Tensor:
        """ATHTDemucs_v2.py:238-248, plus a per-prompt cache (the reference re-runs CLAP per chunk)."""
        missing = [t for t in dict.fromkeys(text) if t not in self._prompt_cache]
        if missing:
            if self.tokenizer is None:
                raise AthtdError(f"no CLAP tokenizer/encoder and no registered embedding for prompts {missing}")
            inputs = self.tokenizer(missing, padding=True, return_tensors="pt")
            inputs = {k: v.to(device) for k, v in inputs.items()}
            with torch.no_grad():
                if hasattr(self.clap, "get_text_features"):
                    e = self.clap.get_text_features(**inputs)
                    e = getattr(e, "pooler_output", e)
                else:
                    e = self.clap.forward(**inputs).text_embeds
            for t, v in zip(missing, e):
                self._prompt_cache[t] = v.detach().float().cpu()
        return torch.stack([self._prompt_cache[t] for t in text]).to(device)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def separate_batch(self, wav: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
        if not wav.is_cuda:
            raise AthtdError("wav must live on the CUDA device (no CPU fallback)")
        B, C, L = wav.shape
        if C != 2:
            raise ValueError("stereo input expected: wav [B,2,T]")
        if emb.dim() == 2:
            emb = emb.unsqueeze(1)
        P = emb.shape[1]
        eng = self.engine(wav.device)
        plan = eng.plan(B, L, P)
        return plan.forward(wav.float().contiguous(), emb.to(wav.device).float().contiguous())

    @torch.no_grad()
    def _encode(self, x: torch.Tensor, xt: torch.Tensor):
        """``AudioTextHTDemucs._encode`` (ATHTDemucs_v2.py:190-236) with the reference's signature and return value:
        x [B,4,2048,Tf] = the normalised complex-as-channels spectrogram, xt [B,2,L] = the normalised waveform ->
        (x_enc [B,384,8,Tf], xt_enc [B,384,St], saved, saved_t, lengths, lengths_t).  ``lengths`` holds the FRAME count
        (the reference stores x.shape[-1] of the 4-D tensor, :198) and ``lengths_t`` the time-branch input lengths (:202).
        The tensors are copies of the workspace buffers converted to float32 in the reference's layouts; forward() does not
        go through this method (its encoder state stays on the device in the kernels' own layout)."""
        if not x.is_cuda:
            raise AthtdError("x must live on the CUDA device (no CPU fallback)")
        B, C4, Fq, Tf = x.shape
        L = xt.shape[-1]
        if C4 != 4 or Fq != 2048 or xt.shape[:2] != (B, 2) or Tf != (L + 1023) // 1024:
            raise ValueError(f"_encode expects x [B,4,2048,ceil(L/1024)] and xt [B,2,L], got {tuple(x.shape)} / {tuple(xt.shape)}")
        from . import lib as _lib
        eng = self.engine(x.device)
        plan = eng.plan(B, L, 1)
        xc = x.float().permute(0, 3, 2, 1).contiguous()                 # [B, Tf, 2048, 4] channels-last
        xtc = xt.float().contiguous()
        with torch.cuda.device(eng.device):
            _lib.check(_lib.load().athtd_encode_normalized(plan.handle, xc.data_ptr(), xtc.data_ptr(),
                                                           torch.cuda.current_stream(eng.device).cuda_stream), "athtd_encode_normalized")
        fr = [512, 128, 32, 8]
        saved, saved_t, lengths, lengths_t = [], [], [], []
        lt = L
        for i in range(4):
            lengths.append(Tf)
            lengths_t.append(lt)
            lt = (lt + 3) // 4
            t = plan.tap(f"enc{i}").interior().float()                   # [B*Tf, F_i, C]
            saved.append(t.view(B, Tf, fr[i], ENC_CH[i]).permute(0, 3, 2, 1).contiguous())
            t = plan.tap(f"tenc{i}").interior().float()                  # [B, L_i, C]
            saved_t.append(t.permute(0, 2, 1).contiguous())
        x_enc = plan.tap("xenc").to_torch().float().view(B, Tf, 8, 384).permute(0, 3, 2, 1).contiguous()
        xt_enc = plan.tap("xtenc").to_torch().float().view(B, -1, 384).permute(0, 2, 1).contiguous()
        return x_enc, xt_enc, saved, saved_t, lengths, lengths_t

    def forward(self, wav: torch.Tensor, text: Union[List[str], str, torch.Tensor]) -> torch.Tensor:
        B = wav.shape[0]
        if isinstance(text, torch.Tensor):
            emb = text
        else:
            if isinstance(text, str):
                text = [text]
            emb = self._get_clap_embeddings(list(text), wav.device)
        if emb.dim() != 2 or emb.shape[0] != B or emb.shape[1] != 512:
            raise ValueError(f"text embeddings must be [B,512] (2-D, one key per segment), got {tuple(emb.shape)}")
        return self.separate_batch(wav, emb)[:, 0]
