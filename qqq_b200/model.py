"""Model-level harness: put `QuantLinear` into a stock Hugging Face decoder.

The reference ships HF subclasses (QQQ/gptq/models/llama.py:165-341, qwen2.py) whose seven decoder linears are
`QuantLinear`, plus `make_quant` / `pack_model` (QQQ/gptq/apply_gptq.py:46-125) that swap `nn.Linear` for
`QuantLinear` by module name and pack on the CPU.  Those subclasses are tied to transformers 4.45 internals
(`LlamaFlashAttention2`, `LlamaSdpaAttention`); here the same result is reached on ANY transformers version by
swapping modules in the stock model:

    find_layers / recurse_setattr     QQQ/utils/model_utils.py:79-89, 112-118
    make_quant(model, names, ...)     QQQ/gptq/apply_gptq.py:91-125          nn.Linear -> empty QuantLinear, by name
    pack_model(model, quantizers,...) QQQ/gptq/apply_gptq.py:46-88           make_quant + QuantLinear.pack per layer
    rtn_quantizers(model, ...)        stand-in for the GPTQ solver (offline calibration is out of scope): round to
                                      nearest with the reference Quantizer's conventions (QQQ/gptq/quant.py:60-93,
                                      QQQ/gptq/gptq.py:204-217), returns the same `quantizers` dict gptq_*_func returns
    build_quantized_model(config, quant_config)   the role of QuantizedLlamaForCausalLM(config, quant_config)
                                      (llama.py:333-341): a model whose decoder linears are empty QuantLinears,
                                      ready for load_state_dict of a QQQ checkpoint
    quantization_config / quantized_state_dict    examples/quant_model.py:322-331, QQQ/utils/utils.py remove_empty_parameters
    fuse_qkv_gate_up(model)           new: q/k/v and gate/up share their input -> ONE activation quant + ONE GEMM
                                      (merge_quant_linears: packed tensors concatenated along N, bit-identical outputs)

Everything between the linears (attention, norms, rotary, SwiGLU) stays stock PyTorch: it is not on the reference's
native hot path either.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

import torch
import torch.nn as nn

from .qlinear import QuantLinear, merge_quant_linears

SUPPORTED_MODEL_TYPES = ("llama", "qwen2")  # QQQ/gptq/models/__init__.py:4-9


def find_layers(module: nn.Module, layers=(nn.Linear,), name: str = "") -> Dict[str, nn.Module]:
    """{qualified name: module} of every submodule whose type is in `layers` (exact type match, like the reference)."""
    if type(module) in tuple(layers):
        return {name: module}
    res = {}
    for child_name, child in module.named_children():
        res.update(find_layers(child, layers, f"{name}.{child_name}" if name else child_name))
    return res


def recurse_setattr(module: nn.Module, name: str, value) -> None:
    head, _, rest = name.partition(".")
    if rest:
        recurse_setattr(getattr(module, head), rest, value)
    else:
        setattr(module, head, value)


def recurse_getattr(module: nn.Module, name: str):
    for part in name.split("."):
        module = getattr(module, part)
    return module


def decoder_linear_names(model: nn.Module) -> Iterable[str]:
    """Names of the linears the reference quantizes: every nn.Linear inside the decoder layers (q,k,v,o,gate,up,down);
    embeddings and lm_head stay fp16 (QQQ/gptq/models/llama.py:307-309, 338-339)."""
    return [n for n in find_layers(model) if ".layers." in f".{n}"]


def make_quant(module: nn.Module, names, bits: int, group_size: int, trainable: bool = False) -> None:
    """Replace every nn.Linear whose qualified name is in `names` by an (unpacked) QuantLinear on the same device."""
    if isinstance(module, QuantLinear):
        return
    names = set(names)
    for name, sub in list(module.named_modules()):
        if name not in names:
            continue
        if not isinstance(sub, nn.Linear):
            raise NotImplementedError(f"{name}: only nn.Linear is swapped (got {type(sub).__name__})")
        dev = sub.weight.device
        new = QuantLinear(bits, group_size, sub.in_features, sub.out_features, sub.bias is not None, trainable=trainable,
                          weight_dtype=sub.weight.dtype)
        new.device = dev
        recurse_setattr(module, name, new.to(dev))


# ---------------------------------------------------------------------------------------------------------------
# round-to-nearest quantizers with the reference Quantizer's conventions
# ---------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def rtn_quantize_weight(W: torch.Tensor, group_size: int, bits: int = 4):
    """W [N, K] -> (W_fq [N, K] float32, scale [N, groups], zero [N, groups], s_extra [N,1] or None).

    per-channel (group_size == -1, symmetric):  maxq = 7,  scale = max|row| / 7, zero = 0,
        W_fq = scale * clamp(round(W/scale), -7, 7)                                  (quant.py:5-13, 36-39, 71-93)
    per-group (symmetric with zero point):      maxq = 15, xmin = min(group, 0), xmax = max(group, 0),
        xmax = max(|xmin|, xmax), xmin = -xmax where xmin < 0, scale = (xmax - xmin) / 15, zero = 8,
        W_fq = scale * (clamp(round(W/scale) + 8, 0, 15) - 8);
        s_extra = max|W_fq row| / 127 — the 8-bit per-channel scale of the fake-quantized weight (gptq.py:204-217).
    All-zero rows/groups get the range [-1, 1] like the reference (quant.py:78-80).
    """
    assert bits == 4
    N, K = W.shape
    W = W.float()
    if group_size == -1 or group_size == K:
        maxq = 2 ** (bits - 1) - 1
        xmax = torch.maximum(W.min(1)[0].clamp(max=0).abs(), W.max(1)[0].clamp(min=0))
        xmax = torch.where(xmax == 0, torch.ones_like(xmax), xmax)
        scale = (xmax / maxq).reshape(N, 1)
        W_fq = scale * torch.clamp(torch.round(W / scale), -maxq, maxq)
        return W_fq, scale, torch.zeros_like(scale), None
    assert K % group_size == 0
    maxq = 2**bits - 1
    Wg = W.reshape(N, K // group_size, group_size)
    xmin = Wg.min(2)[0].clamp(max=0)
    xmax = Wg.max(2)[0].clamp(min=0)
    xmax = torch.maximum(xmin.abs(), xmax)
    xmin = torch.where(xmin < 0, -xmax, xmin)
    both0 = (xmin == 0) & (xmax == 0)
    xmin = torch.where(both0, -torch.ones_like(xmin), xmin)
    xmax = torch.where(both0, torch.ones_like(xmax), xmax)
    scale = (xmax - xmin) / maxq  # [N, groups]
    zero = torch.full_like(scale, (maxq + 1) / 2)
    q = torch.clamp(torch.round(Wg / scale[..., None]) + zero[..., None], 0, maxq)
    W_fq = (scale[..., None] * (q - zero[..., None])).reshape(N, K)
    amax = W_fq.abs().max(1)[0]
    amax = torch.where(amax == 0, torch.ones_like(amax), amax)
    s_extra = (amax / 127.0).reshape(N, 1)
    return W_fq, scale, zero, s_extra


@torch.no_grad()
def rtn_quantizers(model: nn.Module, group_size: int, bits: int = 4, names=None):
    """Fake-quantize the decoder linears IN PLACE (weights replaced by their quantized values, in the weight's dtype)
    and return {name: (scale, zero, g_idx, scale_extra)} — the dict `gptq_llama_func` returns
    (QQQ/gptq/models/llama.py:26-162) and `pack_model` consumes."""
    names = list(names) if names is not None else list(decoder_linear_names(model))
    quantizers = {}
    for name in names:
        lin = recurse_getattr(model, name)
        W_fq, scale, zero, s_extra = rtn_quantize_weight(lin.weight.data, group_size, bits)
        lin.weight.data = W_fq.to(lin.weight.dtype)
        if s_extra is not None:
            # like the reference (gptq.py:191-215) the 8-bit scale is taken from the weight as the layer stores it (after
            # the cast to the layer's dtype); the Quantizer promotes to fp32 (`torch.minimum(x.min(1)[0], zeros_fp32)`,
            # quant.py:70-72), so the maximum of the fp16 values and the division by 127 are fp32 operations
            amax = lin.weight.data.float().abs().max(1)[0]
            s_extra = (torch.where(amax == 0, torch.ones_like(amax), amax) / 127.0).reshape(-1, 1)
        g_idx = torch.arange(lin.in_features, device=scale.device) // (group_size if group_size != -1 else lin.in_features)
        quantizers[name] = (scale, zero, g_idx, s_extra)
    return quantizers


@torch.no_grad()
def pack_model(model: nn.Module, quantizers, bits: int, group_size: int, pack_device: Optional[torch.device] = None):
    """Swap the linears named in `quantizers` for QuantLinears and pack them.  The reference packs on the CPU with numpy
    (apply_gptq.py:70-86); `QuantLinear.pack` here is vectorised torch and runs on whatever device the layer lives on
    (or `pack_device`)."""
    layers = {n: m for n, m in find_layers(model).items() if n in quantizers}
    make_quant(model, quantizers, bits, group_size)
    qlayers = find_layers(model, [QuantLinear])
    for name in quantizers:
        scale, zero, g_idx, s_extra = quantizers[name]
        # the source Linear is released as soon as it is packed (reference: `del layers[name]; free_memory()`,
        # apply_gptq.py:86-87): peak memory is the packed model plus ONE fp16 layer, not plus the whole fp16 model
        ql, lin = qlayers[name], layers.pop(name)
        dev = ql.B.device
        if pack_device is not None:
            ql.to(pack_device)
            lin = lin.to(pack_device)
        pdev = ql.B.device
        ql.pack(lin, scale.to(pdev), s_extra.to(pdev) if s_extra is not None else None)
        ql.to(dev)
        del lin
        if dev.type == "cuda" or (pack_device is not None and torch.device(pack_device).type == "cuda"):
            torch.cuda.empty_cache()
    return model


def quantize_model_rtn(model: nn.Module, group_size: int, bits: int = 4) -> nn.Module:
    """rtn_quantizers + pack_model + the `quantization_config` the reference writes into the HF config."""
    quantizers = rtn_quantizers(model, group_size, bits)
    pack_model(model, quantizers, bits, group_size)
    if hasattr(model, "config"):
        model.config.quantization_config = quantization_config(group_size, bits)
    return model


# ---------------------------------------------------------------------------------------------------------------
# checkpoint format
# ---------------------------------------------------------------------------------------------------------------
def quantization_config(group_size: int, wbits: int = 4) -> dict:
    """examples/quant_model.py:322-327"""
    return {"group_size": group_size, "quant_method": "qqq", "wbits": wbits}


def quantized_state_dict(model: nn.Module) -> dict:
    """State dict as the reference saves it: empty tensors (the per-channel `s_group`) dropped
    (QQQ/utils/utils.py `remove_empty_parameters`); `workspace` / `reduce_buffer` are non-persistent."""
    return {k: v for k, v in model.state_dict().items() if v.numel() > 0}


def get_model_architecture(config) -> str:
    mt = getattr(config, "model_type", None)
    if mt not in SUPPORTED_MODEL_TYPES:
        raise NotImplementedError(f"model_type {mt!r}: the reference supports {SUPPORTED_MODEL_TYPES}")
    return mt


def build_quantized_model(config, quant_config: Optional[dict] = None, dtype=torch.float16, device=None) -> nn.Module:
    """A causal LM of `config` whose decoder linears are (empty) QuantLinears — the role of
    `get_quantized_model_class(model_type)(config, quant_config)` (QQQ/gptq/models/__init__.py:18-22,
    llama.py:333-341).  Load a QQQ checkpoint into it with `load_quantized_state_dict`."""
    from transformers import AutoModelForCausalLM

    get_model_architecture(config)
    quant_config = quant_config or getattr(config, "quantization_config", None)
    if not isinstance(quant_config, dict):
        quant_config = quant_config.to_dict() if hasattr(quant_config, "to_dict") else dict(quant_config or {})
    if quant_config.get("quant_method", "qqq") != "qqq":
        raise ValueError(f"quant_method {quant_config.get('quant_method')!r} is not 'qqq'")
    bits, group_size = int(quant_config.get("wbits", 4)), int(quant_config.get("group_size", -1))
    # the HF config must not carry an unknown quantization_config into from_config (HF would look for a quantizer)
    cfg = config.__class__.from_dict({k: v for k, v in config.to_dict().items() if k != "quantization_config"})
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        if device is not None:
            with torch.device(device):
                model = AutoModelForCausalLM.from_config(cfg)
        else:
            model = AutoModelForCausalLM.from_config(cfg)
    finally:
        torch.set_default_dtype(prev)
    make_quant(model, decoder_linear_names(model), bits, group_size)
    model.config.quantization_config = quantization_config(group_size, bits)
    return model


def load_quantized_state_dict(model: nn.Module, state_dict: dict) -> None:
    """Strict load, except that the empty per-channel `s_group` buffers (dropped by `quantized_state_dict`) may be absent."""
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    missing = [k for k in missing if not (k.endswith(".s_group") and recurse_getattr(model, k).numel() == 0)]
    if missing or unexpected:
        raise RuntimeError(f"load_quantized_state_dict: missing keys {missing}, unexpected keys {list(unexpected)}")


# ---------------------------------------------------------------------------------------------------------------
# q/k/v and gate/up fusion
# ---------------------------------------------------------------------------------------------------------------
class _SharedGemm(nn.Module):
    """One merged QuantLinear serving several projection slots that consume the same input tensor."""

    def __init__(self, merged: QuantLinear):
        super().__init__()
        self.merged = merged
        self._key = None
        self._ref = None  # keeps the keyed input alive: its address cannot be handed to another tensor while we hold it
        self._out = None
        self._served = 0

    @staticmethod
    def _key_of(x: torch.Tensor):
        # inference tensors have no version counter (reading `_version` raises); for them identity of the tensor object,
        # which `_ref` keeps alive, stands in for it
        ver = id(x) if x.is_inference() else x._version
        return (x.untyped_storage().data_ptr(), x.storage_offset(), tuple(x.shape), tuple(x.stride()), x.dtype, ver, x.device)

    def slice(self, x: torch.Tensor, index: int) -> torch.Tensor:
        key = self._key_of(x)
        # slot 0 (q_proj / gate_proj) is always called first in a forward: it always recomputes, so an entry left behind by an
        # interrupted forward (an exception between q_proj and v_proj) can never be served to the next one
        if index == 0 or self._key != key or self._out is None:
            self._key, self._ref, self._out = None, None, None
            self._out = self.merged(x).split(self.merged.split_sizes, dim=-1)
            self._key, self._ref = key, x
            self._served = 0
        y = self._out[index]
        self._served += 1
        if self._served == len(self.merged.split_sizes):  # every consumer has its slice: drop the references
            self._key, self._ref, self._out = None, None, None
        return y


class FusedProjection(nn.Module):
    """Stands where `q_proj` / `k_proj` / `v_proj` (or `gate_proj` / `up_proj`) stood: returns its column slice of the
    shared merged GEMM.  The first slot called with a new input runs the GEMM; the others reuse its output."""

    def __init__(self, shared: _SharedGemm, index: int, owner: bool):
        super().__init__()
        if owner:
            self.shared = shared  # registered once, so the merged buffers appear once in the module tree
        else:
            object.__setattr__(self, "_shared_ref", shared)
        self.index = index
        self.in_features = shared.merged.infeatures
        self.out_features = shared.merged.split_sizes[index]

    def _shared(self) -> _SharedGemm:
        return self.shared if "shared" in self._modules else self._shared_ref

    def forward(self, x):
        return self._shared().slice(x, self.index)


_FUSE_GROUPS = (("self_attn", ("q_proj", "k_proj", "v_proj")), ("mlp", ("gate_proj", "up_proj")))


def fuse_qkv_gate_up(model: nn.Module) -> int:
    """Merge q/k/v and gate/up QuantLinears of every decoder layer (HF Llama / Qwen2 layout).  Returns the number of
    merged groups.  Outputs are bit-identical to the unmerged model (tests/test_model_harness.py)."""
    n = 0
    for layer in _decoder_layers(model):
        for parent_name, slots in _FUSE_GROUPS:
            parent = getattr(layer, parent_name, None)
            if parent is None:
                continue
            mods = [getattr(parent, s, None) for s in slots]
            if not all(isinstance(m, QuantLinear) for m in mods):
                continue
            if any(m.outfeatures % 64 for m in mods):
                continue
            shared = _SharedGemm(merge_quant_linears(mods))
            for i, s in enumerate(slots):
                setattr(parent, s, FusedProjection(shared, i, owner=(i == 0)))
            n += 1
    return n


def share_scratch(model: nn.Module) -> int:
    """Point every QuantLinear of `model` at ONE split-K scratch and ONE lock array per device instead of its own
    `reduce_buffer` (64*max_par x N int32 — 5.5 GB summed over Llama-2-7B) and `workspace`.  Safe because launches of one
    model are ordered on a stream, the kernel touches scratch and locks only after the preceding kernel has completed
    (griddepcontrol.wait) and returns the locks zeroed.  Each module keeps a [64*max_par, N] VIEW, so `qqq_gemm` still
    reads N from `C.size(1)` like the reference (csrc/qqq_gemm.cu:1063).  Returns the bytes released."""
    by_dev = {}
    for ql in find_layers(model, [QuantLinear]).values():
        by_dev.setdefault(ql.B.device, []).append(ql)
    freed = 0
    for dev, qls in by_dev.items():
        rows = max(q.max_par * 64 for q in qls)
        flat = torch.zeros(rows * max(q.outfeatures for q in qls), dtype=torch.int32, device=dev)
        locks = torch.zeros(max(max(q.outfeatures // 128 * q.max_par, 16) for q in qls), dtype=torch.int32, device=dev)
        for q in qls:
            freed += q.reduce_buffer.numel() * 4 + q.workspace.numel() * 4
            q.reduce_buffer = flat[: q.max_par * 64 * q.outfeatures].view(q.max_par * 64, q.outfeatures)
            q.workspace = locks
        freed -= flat.numel() * 4 + locks.numel() * 4
    return freed


def _decoder_layers(model: nn.Module):
    for cand in ("model.layers", "layers", "model.model.layers"):
        try:
            return list(recurse_getattr(model, cand))
        except AttributeError:
            continue
    raise AttributeError("no decoder layer list found (expected `model.layers`)")


__all__ = ["find_layers", "recurse_setattr", "recurse_getattr", "decoder_linear_names", "make_quant", "rtn_quantize_weight",
           "rtn_quantizers", "pack_model", "quantize_model_rtn", "quantization_config", "quantized_state_dict",
           "get_model_architecture", "build_quantized_model", "load_quantized_state_dict", "fuse_qkv_gate_up", "share_scratch",
           "FusedProjection"]
