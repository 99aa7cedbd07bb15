"""CLIP relevancy engine façade: ``interpret()`` with the reference signature (CLIP_explainability.ipynb cell 6)
on top of ``mmx_clip_*`` (include/mmx.h).  PyTorch only owns memory and streams here."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Tuple

import torch

from ._lib import lib, check, ptr, current_stream, ClipConfigC, MmxError


@dataclass(frozen=True)
class ClipConfig:
    """Constructor arguments of the reference ``CLIP`` (CLIP/clip/model.py:249-262), ViT image tower."""
    embed_dim: int = 512
    image_resolution: int = 224
    vision_layers: int = 12
    vision_width: int = 768
    vision_patch_size: int = 32
    context_length: int = 77
    vocab_size: int = 49408
    transformer_width: int = 512
    transformer_heads: int = 8
    transformer_layers: int = 12

    @property
    def vision_tokens(self) -> int:
        return (self.image_resolution // self.vision_patch_size) ** 2 + 1

    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor]) -> "ClipConfig":
        """Infers every dimension from the weights exactly like ``build_model`` (CLIP/clip/model.py:405-434)."""
        if "visual.proj" not in sd:
            raise MmxError("only ViT image towers are supported (ModifiedResNet CLIP is out of the hot-path scope)")
        vision_width = sd["visual.conv1.weight"].shape[0]
        vision_layers = len([k for k in sd if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
        patch = sd["visual.conv1.weight"].shape[-1]
        grid = round((sd["visual.positional_embedding"].shape[0] - 1) ** 0.5)
        width = sd["ln_final.weight"].shape[0]
        layers = len(set(k.split(".")[2] for k in sd if k.startswith("transformer.resblocks")))
        return cls(sd["text_projection"].shape[1], patch * grid, vision_layers, vision_width, patch,
                   sd["positional_embedding"].shape[0], sd["token_embedding.weight"].shape[0], width, width // 64, layers)

    def to_c(self) -> ClipConfigC:
        return ClipConfigC(self.embed_dim, self.image_resolution, self.vision_layers, self.vision_width,
                           self.vision_patch_size, self.context_length, self.vocab_size, self.transformer_width,
                           self.transformer_heads, self.transformer_layers)


VIT_B32 = ClipConfig()
VIT_L14_336 = ClipConfig(768, 336, 24, 1024, 14, 77, 49408, 768, 12, 12)


class ClipEngine:
    """Device-resident CLIP relevancy engine.  Plays the role of the reference's ``model`` argument of
    ``interpret()``; build it from a reference ``state_dict`` (``ClipEngine.from_state_dict(model.state_dict())``)."""

    def __init__(self, cfg: ClipConfig, state_dict: Dict[str, torch.Tensor], max_batch: int = 64, device=None):
        if not torch.cuda.is_available():
            raise MmxError("mmx_b200 needs a CUDA (sm_100) device; there is no CPU fallback")
        self.cfg = cfg
        self.max_batch = int(max_batch)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._h = C.c_void_p()
        self._lib = lib()
        with torch.cuda.device(self.device):
            ccfg = cfg.to_c()
            check(self._lib.mmx_clip_create(C.byref(ccfg), self.max_batch, C.byref(self._h)))
            for name, t in state_dict.items():
                if name in ("input_resolution", "context_length", "vocab_size") or name.endswith("attn_mask"):
                    continue   # bookkeeping entries of JIT archives (CLIP/clip/model.py:437-439)
                t = t.detach().to("cpu", torch.float32).contiguous()
                check(self._lib.mmx_clip_load_tensor(self._h, name.encode(), C.c_void_p(t.data_ptr()), t.numel()))
            check(self._lib.mmx_clip_finalize(self._h))

    @classmethod
    def from_state_dict(cls, sd, max_batch: int = 64, device=None) -> "ClipEngine":
        return cls(ClipConfig.from_state_dict(sd), sd, max_batch, device)

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                self._lib.mmx_clip_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def eval(self):
        return self

    def zero_grad(self):
        return None

    # -- device path: inputs are CUDA tensors, work is enqueued on torch's current stream
    def interpret(self, images: torch.Tensor, tokens: torch.Tensor, start_layer: int = -1, start_layer_text: int = -1
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
        B = tokens.shape[0]
        n_img = images.shape[0]
        if n_img not in (1, B):
            raise MmxError("images must hold 1 (repeated, as in the notebook) or B images")
        with torch.cuda.device(self.device):
            images = images.to(self.device, torch.float32).contiguous()
            tok = tokens.to(self.device, torch.int32).contiguous()
            ctx, sv = self.cfg.context_length, self.cfg.vision_tokens
            R_text = torch.empty(B, ctx, ctx, device=self.device, dtype=torch.float32)
            R_image = torch.empty(B, sv - 1, device=self.device, dtype=torch.float32)
            check(self._lib.mmx_clip_interpret_device(self._h, ptr(images), n_img, ptr(tok), B, int(start_layer),
                                                      int(start_layer_text), ptr(R_text), ptr(R_image), current_stream()))
        return R_text, R_image

    # -- host path: numpy / CPU tensors in, CPU tensors out; copies are inside the C call
    def interpret_host(self, images: torch.Tensor, tokens: torch.Tensor, start_layer: int = -1, start_layer_text: int = -1,
                       out: Tuple[torch.Tensor, torch.Tensor] | None = None):
        B, n_img = tokens.shape[0], images.shape[0]
        assert not images.is_cuda and not tokens.is_cuda
        images = images.to(torch.float32).contiguous()
        tok = tokens.to(torch.int32).contiguous()
        ctx, sv = self.cfg.context_length, self.cfg.vision_tokens
        if out is None:
            out = (torch.empty(B, ctx, ctx, dtype=torch.float32).pin_memory(),
                   torch.empty(B, sv - 1, dtype=torch.float32).pin_memory())
        with torch.cuda.device(self.device):
            check(self._lib.mmx_clip_interpret_host(self._h, ptr(images), n_img, ptr(tok), B, int(start_layer),
                                                    int(start_layer_text), ptr(out[0]), ptr(out[1])))
        return out

    def tap(self, what: str, tower: int = 0, layer: int = 0) -> torch.Tensor:
        """Copy of an intermediate of the last interpret call: 'A', 'dA' [B,H,S,S], 'Abar' [B,S,S], 'logits' [B,B]."""
        p = C.c_void_p()
        dims = (C.c_int * 4)()
        ld = C.c_int()
        check(self._lib.mmx_clip_tap(self._h, what.encode(), tower, layer, C.byref(p), C.byref(dims), C.byref(ld)))
        shape = [d for d in dims]
        n_lead = 1
        nd = {"A": 4, "dA": 4, "Abar": 3, "logits": 2}[what]
        shape = shape[:nd]
        for d in shape[:-1]:
            n_lead *= d
        flat = torch.empty(n_lead * ld.value, device=self.device, dtype=torch.float32)
        torch.cuda.synchronize(self.device)
        check(self._lib.mmx_memcpy_d2d(ptr(flat), p, flat.numel() * 4, current_stream()))
        torch.cuda.synchronize(self.device)
        return flat.view(n_lead, ld.value)[:, :shape[-1]].reshape(shape).clone()


def interpret(image, texts, model, device=None, start_layer=-1, start_layer_text=-1):
    """Drop-in for the notebook's ``interpret`` (CLIP_explainability.ipynb:151-208).

    ``image``: [1,3,R,R] (repeated for every text, :153) or [B,3,R,R] (one image per text); ``texts``: [B,ctx] token
    ids; ``model``: a :class:`ClipEngine`.  Returns ``(text_relevance [B,ctx,ctx], image_relevance [B,S-1])`` on the
    engine's device."""
    if not isinstance(model, ClipEngine):
        raise MmxError("model must be a mmx_b200.ClipEngine (build one with ClipEngine.from_state_dict(ref_model.state_dict()))")
    return model.interpret(image, texts, start_layer, start_layer_text)
