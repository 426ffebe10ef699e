"""ctypes binding of include/fseend_b200.h.  torch is used only for device memory and streams.

There is NO CPU fallback: ``lib()`` raises if the shared library has not been built, and model creation
raises if the current device is not an sm_100 GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .build import LIB_PATH

_lib = None


class FseendError(RuntimeError):
    pass


class FsConfig(C.Structure):
    _fields_ = [
        ("in_size", C.c_int), ("n_units", C.c_int), ("n_heads", C.c_int), ("enc_n_layers", C.c_int),
        ("dec_n_layers", C.c_int), ("enc_dim_feedforward", C.c_int), ("dec_dim_feedforward", C.c_int),
        ("conv_kernel", C.c_int), ("conv_padding", C.c_int), ("mask_delay", C.c_int), ("has_mask", C.c_int),
        ("bn_eps", C.c_float), ("ln_eps", C.c_float),
    ]


class LsConfig(C.Structure):
    _fields_ = [(k, C.c_int) for k in (
        "in_size", "n_units", "n_heads", "enc_n_layers", "dec_n_layers", "feed_forward_expansion_factor",
        "dec_dim_feedforward", "conv_kernel_size", "recurrent_chunk_size", "conv_delay")]


# every symbol include/fseend_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "fseend_version", "fseend_last_error", "fseend_device_ok", "fseend_fs_create", "fseend_fs_destroy",
    "fseend_fs_forward", "fseend_fs_forward_host", "fseend_fs_set_profiling", "fseend_fs_get_profile",
    "fseend_fs_launches_per_forward", "fseend_fs_workspace_bytes", "fseend_fs_set_option", "fseend_fs_stream_create",
    "fseend_fs_stream_destroy", "fseend_fs_stream_step", "fseend_fs_stream_frames", "fseend_ls_create",
    "fseend_ls_destroy", "fseend_ls_padded_len", "fseend_ls_forward", "fseend_ls_launches_per_forward",
    "fseend_ls_stream_create", "fseend_ls_stream_destroy", "fseend_ls_stream_reset", "fseend_ls_stream_step",
    "fseend_ls_stream_enc_step", "fseend_ls_stream_dec_step", "fseend_op_gemm",
    "fseend_op_gemm_ex", "fseend_op_retention", "fseend_op_dwconv_bn_swish", "fseend_op_ret_step",
    "fseend_op_ffn", "fseend_op_causal_attn", "fseend_op_spk_attn", "fseend_op_spk_attn_tc", "fseend_op_head",
    "fseend_op_prep_input", "fseend_op_embloss", "fseend_op_embloss_workspace_bytes", "fseend_op_spk_qkv_attn", "fseend_op_decide_median", "fseend_op_label_prepare", "fseend_op_splice_subsample",
    "fseend_op_bce_loss", "fseend_op_bce_loss_workspace_bytes",
    "fseend_fs_forward_host_async", "fseend_fs_host_wait", "fseend_op_pit_costs", "fseend_ls_forward_host", "fseend_ls_set_option", "fseend_ls_get_option", "fseend_p32_linear_create",
    "fseend_p32_linear_destroy", "fseend_p32_linear_apply", "fseend_op_p32_retention",
    "fseend_train_linear_workspace_bytes", "fseend_train_linear_fwd", "fseend_train_linear_bwd", "fseend_train_add_layernorm_fwd",
    "fseend_train_layernorm_workspace_bytes", "fseend_train_layernorm_bwd", "fseend_train_attn_fwd", "fseend_train_attn_bwd",
    "fseend_train_spk_attn_fwd", "fseend_train_spk_attn_bwd", "fseend_train_l2norm_fwd", "fseend_train_l2norm_bwd",
    "fseend_train_head_fwd", "fseend_train_head_bwd", "fseend_train_batchnorm_workspace_bytes", "fseend_train_batchnorm_fwd",
    "fseend_train_batchnorm_bwd",
]


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FseendError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  fseend_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, ip, fp = C.c_void_p, C.c_int, C.c_float
    L.fseend_version.restype = ip
    L.fseend_last_error.restype = C.c_char_p
    L.fseend_device_ok.restype = ip
    L.fseend_fs_create.restype = ip
    L.fseend_fs_create.argtypes = [C.POINTER(FsConfig), ip, C.POINTER(C.c_char_p), C.POINTER(vp),
                                   C.POINTER(C.c_longlong), C.POINTER(vp)]
    L.fseend_fs_destroy.restype = None
    L.fseend_fs_destroy.argtypes = [vp]
    L.fseend_fs_forward.restype = ip
    L.fseend_fs_forward.argtypes = [vp, vp, C.POINTER(ip), ip, ip, vp, vp, vp, vp]
    L.fseend_fs_forward_host.restype = ip
    L.fseend_fs_forward_host.argtypes = [vp, vp, C.POINTER(ip), ip, ip, vp, vp, vp]
    L.fseend_fs_forward_host_async.restype = ip
    L.fseend_fs_forward_host_async.argtypes = [vp, vp, C.POINTER(ip), ip, ip, vp, C.POINTER(C.c_longlong)]
    L.fseend_fs_host_wait.restype = ip
    L.fseend_fs_host_wait.argtypes = [vp, C.c_longlong]
    L.fseend_fs_set_profiling.restype = ip
    L.fseend_fs_set_profiling.argtypes = [vp, ip]
    L.fseend_fs_get_profile.restype = ip
    L.fseend_fs_get_profile.argtypes = [vp, ip, vp, C.POINTER(fp), C.POINTER(ip)]
    L.fseend_fs_launches_per_forward.restype = ip
    L.fseend_fs_launches_per_forward.argtypes = [vp]
    L.fseend_fs_workspace_bytes.restype = C.c_size_t
    L.fseend_fs_workspace_bytes.argtypes = [vp]
    L.fseend_op_gemm.restype = ip
    L.fseend_op_gemm.argtypes = [vp, ip, ip, ip, vp, ip, ip, ip, ip, ip, vp, vp, vp, vp, fp, vp, ip, vp, vp, vp]
    L.fseend_op_causal_attn.restype = ip
    L.fseend_op_causal_attn.argtypes = [vp, ip, ip, ip, ip, ip, fp, vp, vp]
    L.fseend_fs_stream_create.restype = ip
    L.fseend_fs_stream_create.argtypes = [vp, ip, ip, C.POINTER(vp)]
    L.fseend_fs_stream_destroy.restype = None
    L.fseend_fs_stream_destroy.argtypes = [vp]
    L.fseend_fs_stream_step.restype = ip
    L.fseend_fs_stream_step.argtypes = [vp, vp, vp, C.POINTER(ip), vp]
    L.fseend_fs_stream_frames.restype = ip
    L.fseend_fs_stream_frames.argtypes = [vp]
    L.fseend_ls_create.restype = ip
    L.fseend_ls_create.argtypes = [C.POINTER(LsConfig), ip, C.POINTER(C.c_char_p), C.POINTER(vp),
                                   C.POINTER(C.c_longlong), C.POINTER(vp)]
    L.fseend_ls_destroy.restype = None
    L.fseend_ls_destroy.argtypes = [vp]
    L.fseend_ls_padded_len.restype = ip
    L.fseend_ls_padded_len.argtypes = [vp, ip]
    L.fseend_ls_forward.restype = ip
    L.fseend_ls_forward.argtypes = [vp, vp, C.POINTER(ip), ip, ip, vp, vp, vp, vp]
    L.fseend_ls_launches_per_forward.restype = ip
    L.fseend_ls_launches_per_forward.argtypes = [vp]
    L.fseend_ls_forward_host.restype = ip
    L.fseend_ls_forward_host.argtypes = [vp, vp, C.POINTER(ip), ip, ip, vp, vp, vp]
    L.fseend_ls_set_option.restype = ip
    L.fseend_ls_set_option.argtypes = [vp, C.c_char_p, ip]
    L.fseend_ls_get_option.restype = ip
    L.fseend_ls_get_option.argtypes = [vp, C.c_char_p]
    L.fseend_p32_linear_create.restype = ip
    L.fseend_p32_linear_create.argtypes = [vp, ip, ip, C.POINTER(vp)]
    L.fseend_p32_linear_destroy.restype = None
    L.fseend_p32_linear_destroy.argtypes = [vp]
    L.fseend_p32_linear_apply.restype = ip
    L.fseend_p32_linear_apply.argtypes = [vp, vp, ip, vp, ip, fp, vp, vp, vp]
    L.fseend_op_p32_retention.restype = ip
    L.fseend_op_p32_retention.argtypes = [vp, ip, ip, ip, ip, vp, vp]
    L.fseend_ls_stream_create.restype = ip
    L.fseend_ls_stream_create.argtypes = [vp, ip, ip, C.POINTER(vp)]
    L.fseend_ls_stream_destroy.restype = None
    L.fseend_ls_stream_destroy.argtypes = [vp]
    L.fseend_ls_stream_reset.restype = ip
    L.fseend_ls_stream_reset.argtypes = [vp]
    L.fseend_ls_stream_step.restype = ip
    L.fseend_ls_stream_step.argtypes = [vp, vp, vp, C.POINTER(ip), vp]
    L.fseend_ls_stream_enc_step.restype = ip
    L.fseend_ls_stream_enc_step.argtypes = [vp, vp, ip, vp, vp]
    L.fseend_ls_stream_dec_step.restype = ip
    L.fseend_ls_stream_dec_step.argtypes = [vp, vp, ip, vp, vp]
    L.fseend_op_gemm_ex.restype = ip
    L.fseend_op_gemm_ex.argtypes = [vp, ip, ip, ip, vp, ip, ip, ip, vp, vp, fp, vp, vp, vp, vp, fp, vp, vp, vp, vp]
    L.fseend_op_retention.restype = ip
    L.fseend_op_retention.argtypes = [vp, ip, ip, ip, ip, vp, vp, vp, vp]
    L.fseend_op_dwconv_bn_swish.restype = ip
    L.fseend_op_dwconv_bn_swish.argtypes = [vp, vp, vp, vp, ip, ip, ip, vp, vp, vp]
    L.fseend_op_ret_step.restype = ip
    L.fseend_op_ret_step.argtypes = [vp, vp, ip, ip, vp, vp]
    L.fseend_fs_set_option.restype = ip
    L.fseend_fs_set_option.argtypes = [vp, C.c_char_p, ip]
    L.fseend_op_ffn.restype = ip
    L.fseend_op_ffn.argtypes = [vp, ip, ip, vp, vp, vp, vp, ip, vp, vp, fp, vp, ip, vp, vp]
    L.fseend_op_spk_attn_tc.restype = ip
    L.fseend_op_spk_attn_tc.argtypes = [vp, ip, ip, fp, vp, vp]
    L.fseend_op_spk_attn.restype = ip
    L.fseend_op_spk_attn.argtypes = [vp, ip, ip, fp, vp, vp]
    L.fseend_op_head.restype = ip
    L.fseend_op_head.argtypes = [vp, vp, ip, ip, vp, vp, vp, vp]
    L.fseend_op_prep_input.restype = ip
    L.fseend_op_prep_input.argtypes = [vp, vp, ip, ip, ip, ip, vp, vp, vp, vp]
    L.fseend_op_spk_qkv_attn.restype = ip
    L.fseend_op_spk_qkv_attn.argtypes = [vp, vp, vp, ip, ip, fp, vp, vp]
    L.fseend_op_label_prepare.restype = ip
    L.fseend_op_label_prepare.argtypes = [vp, ip, ip, ip, vp, vp, vp]
    L.fseend_op_bce_loss_workspace_bytes.restype = C.c_size_t
    L.fseend_op_bce_loss_workspace_bytes.argtypes = [ip, ip]
    L.fseend_op_bce_loss.restype = ip
    L.fseend_op_bce_loss.argtypes = [vp, ip, vp, ip, ip, ip, vp, vp, ip, vp, vp, vp]
    L.fseend_op_pit_costs.restype = ip
    L.fseend_op_pit_costs.argtypes = [vp, vp, ip, ip, ip, vp, ip, ip, vp, vp]
    L.fseend_op_splice_subsample.restype = ip
    L.fseend_op_splice_subsample.argtypes = [vp, ip, ip, ip, ip, vp, vp]
    L.fseend_op_decide_median.restype = ip
    L.fseend_op_decide_median.argtypes = [vp, ip, ip, fp, ip, vp, vp]
    L.fseend_op_embloss_workspace_bytes.restype = C.c_size_t
    L.fseend_op_embloss_workspace_bytes.argtypes = [ip, ip]
    L.fseend_op_embloss.restype = ip
    L.fseend_op_embloss.argtypes = [vp, vp, vp, ip, ip, ip, C.c_double, vp, vp, vp]
    sz = C.c_size_t
    L.fseend_train_linear_workspace_bytes.restype = sz
    L.fseend_train_linear_workspace_bytes.argtypes = [ip, ip, ip]
    L.fseend_train_linear_fwd.restype = ip
    L.fseend_train_linear_fwd.argtypes = [vp, ip, ip, vp, ip, vp, ip, vp, vp, sz, vp]
    L.fseend_train_linear_bwd.restype = ip
    L.fseend_train_linear_bwd.argtypes = [vp, vp, vp, vp, ip, ip, ip, ip, ip, vp, vp, vp, vp, sz, vp]
    L.fseend_train_add_layernorm_fwd.restype = ip
    L.fseend_train_add_layernorm_fwd.argtypes = [vp, vp, vp, vp, ip, fp, vp, vp, vp]
    L.fseend_train_layernorm_workspace_bytes.restype = sz
    L.fseend_train_layernorm_workspace_bytes.argtypes = [ip]
    L.fseend_train_layernorm_bwd.restype = ip
    L.fseend_train_layernorm_bwd.argtypes = [vp, vp, vp, ip, fp, vp, vp, vp, vp, sz, vp]
    L.fseend_train_attn_fwd.restype = ip
    L.fseend_train_attn_fwd.argtypes = [vp, ip, ip, ip, ip, fp, C.c_ulonglong, vp, vp, vp]
    L.fseend_train_attn_bwd.restype = ip
    L.fseend_train_attn_bwd.argtypes = [vp, vp, vp, vp, ip, ip, ip, ip, fp, C.c_ulonglong, vp, vp, vp]
    L.fseend_train_spk_attn_fwd.restype = ip
    L.fseend_train_spk_attn_fwd.argtypes = [vp, ip, ip, fp, C.c_ulonglong, vp, vp]
    L.fseend_train_spk_attn_bwd.restype = ip
    L.fseend_train_spk_attn_bwd.argtypes = [vp, vp, ip, ip, fp, C.c_ulonglong, vp, vp]
    L.fseend_train_l2norm_fwd.restype = ip
    L.fseend_train_l2norm_fwd.argtypes = [vp, ip, vp, vp, vp]
    L.fseend_train_l2norm_bwd.restype = ip
    L.fseend_train_l2norm_bwd.argtypes = [vp, vp, vp, ip, vp, vp]
    L.fseend_train_head_fwd.restype = ip
    L.fseend_train_head_fwd.argtypes = [vp, vp, ip, ip, vp, vp]
    L.fseend_train_head_bwd.restype = ip
    L.fseend_train_head_bwd.argtypes = [vp, vp, vp, ip, ip, vp, vp, vp]
    L.fseend_train_batchnorm_workspace_bytes.restype = sz
    L.fseend_train_batchnorm_workspace_bytes.argtypes = [ip, ip]
    L.fseend_train_batchnorm_fwd.restype = ip
    L.fseend_train_batchnorm_fwd.argtypes = [vp, vp, vp, ip, ip, fp, vp, vp, vp, sz, vp]
    L.fseend_train_batchnorm_bwd.restype = ip
    L.fseend_train_batchnorm_bwd.argtypes = [vp, vp, vp, vp, ip, ip, fp, vp, vp, vp, vp, sz, vp]
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise FseendError(f"fseend error {rc}: {lib().fseend_last_error().decode()}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*ts: torch.Tensor):
    for t in ts:
        if t is not None and (not t.is_cuda or not t.is_contiguous()):
            raise FseendError("fseend_b200 kernels take contiguous CUDA tensors (no CPU fallback)")


class NativeCacheMixin:
    """Lazy native model of an nn.Module that owns reference-named parameters.

    The cache key is (data_ptr, _version) of every parameter/buffer plus the device they live on.  In-place edits through
    ``.data`` (``p.data.copy_(...)``, the style of the reference's own copy_params) do NOT bump ``_version``: after such an
    edit call ``invalidate_native()``.  ``load_state_dict``, optimizer steps and ``.to()`` are picked up automatically.
    The native model is created on the device the parameters live on, not on ``torch.cuda.current_device()``."""

    _native = None
    _native_key = None

    def invalidate_native(self):
        """Drop the cached native model (and any streaming state built on it); the next call rebuilds it."""
        self._native = None
        self._native_key = None
        if hasattr(self, "_stream"):
            self._stream = None

    def _on_device(self):
        """Context that makes the model's GPU the current device for a native call (kernels launch on the current
        device; workspaces and weights live on the model's)."""
        return torch.cuda.device(next(self.parameters()).device)

    def _native_cached(self, build):
        tensors = list(self.parameters()) + list(self.buffers())
        dev = tensors[0].device
        if dev.type != "cuda":
            raise FseendError("fseend_b200 runs on a CUDA sm_100 device only (move the model with .cuda())")
        key = (tuple((t.data_ptr(), t._version) for t in tensors), dev.index)
        if self._native is None or key != self._native_key:
            with torch.cuda.device(dev):
                self._native = build()
            self._native_key = key
            if hasattr(self, "_stream"):
                self._stream = None
        return self._native


class FsModel:
    """Owns a native fseend_fs_model built from a reference-named state_dict."""

    def __init__(self, cfg: Dict[str, int], state_dict: Dict[str, torch.Tensor]):
        L = lib()
        if not torch.cuda.is_available():
            raise FseendError("fseend_b200 requires a CUDA (sm_100) device; there is no CPU fallback")
        c = FsConfig(
            in_size=cfg["in_size"], n_units=cfg["n_units"], n_heads=cfg["n_heads"],
            enc_n_layers=cfg["enc_n_layers"], dec_n_layers=cfg["dec_n_layers"],
            enc_dim_feedforward=cfg.get("enc_dim_feedforward", 2048),
            dec_dim_feedforward=cfg["dec_dim_feedforward"], conv_kernel=cfg.get("conv_kernel", 19),
            conv_padding=cfg.get("conv_padding", 9), mask_delay=cfg.get("mask_delay", 0), has_mask=int(cfg.get("has_mask", True)),
            bn_eps=cfg.get("bn_eps", 1e-5), ln_eps=cfg.get("ln_eps", 1e-5))
        self.cfg = dict(cfg)
        names, ptrs, numels, keep = [], [], [], []
        for k, v in state_dict.items():
            if not torch.is_floating_point(v):
                continue
            t = v.detach().to(device="cpu", dtype=torch.float32).contiguous()
            keep.append(t)
            names.append(k.encode())
            ptrs.append(t.data_ptr())
            numels.append(t.numel())
        n = len(names)
        handle = C.c_void_p()
        _check(L.fseend_fs_create(C.byref(c), n, (C.c_char_p * n)(*names), (C.c_void_p * n)(*ptrs),
                                  (C.c_longlong * n)(*numels), C.byref(handle)))
        self._h = handle
        self._L = L

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.fseend_fs_destroy(h)

    def forward(self, x_packed: torch.Tensor, ilens: Sequence[int], max_nspks: int, want_emb: bool = False,
                want_att: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
        """x_packed: CUDA fp32 [sum(ilens), in_size].  Returns padded (logits [B,T,S], emb [B,T,D], att [B,T,S,D])."""
        _require_cuda(x_packed)
        if x_packed.dtype != torch.float32:
            raise FseendError("x must be float32")
        B, T = len(ilens), int(max(ilens))
        if x_packed.shape[0] != int(sum(ilens)) or x_packed.shape[1] != self.cfg["in_size"]:
            raise FseendError("x_packed shape does not match ilens / in_size")
        D = self.cfg["n_units"]
        dev = x_packed.device
        logits = torch.empty(B, T, max_nspks, device=dev, dtype=torch.float32)
        emb = torch.empty(B, T, D, device=dev, dtype=torch.float32) if want_emb else None
        att = torch.empty(B, T, max_nspks, D, device=dev, dtype=torch.float32) if want_att else None
        il = (C.c_int * B)(*[int(i) for i in ilens])
        _check(self._L.fseend_fs_forward(self._h, _ptr(x_packed), il, B, max_nspks, _ptr(logits), _ptr(emb), _ptr(att),
                                         _stream()))
        return logits, emb, att

    def forward_host(self, x_packed: torch.Tensor, ilens: Sequence[int], max_nspks: int, want_emb: bool = False,
                     want_att: bool = False, out: Optional[torch.Tensor] = None):
        """Host-buffer entry point (H2D + forward + D2H inside the call).  x_packed: CPU fp32 (pinned or not)."""
        if x_packed.is_cuda or x_packed.dtype != torch.float32 or not x_packed.is_contiguous():
            raise FseendError("forward_host takes a contiguous CPU float32 tensor")
        B, T = len(ilens), int(max(ilens))
        D = self.cfg["n_units"]
        logits = out if out is not None else torch.empty(B, T, max_nspks, dtype=torch.float32)
        emb = torch.empty(B, T, D, dtype=torch.float32) if want_emb else None
        att = torch.empty(B, T, max_nspks, D, dtype=torch.float32) if want_att else None
        il = (C.c_int * B)(*[int(i) for i in ilens])
        _check(self._L.fseend_fs_forward_host(self._h, _ptr(x_packed), il, B, max_nspks, _ptr(logits), _ptr(emb),
                                              _ptr(att)))
        return logits, emb, att

    def forward_host_async(self, x_packed: torch.Tensor, ilens: Sequence[int], max_nspks: int, out: torch.Tensor) -> int:
        """Pipelined host-buffer entry point: enqueues H2D + forward + D2H and returns a ticket; ``host_wait(ticket)``
        blocks until ``out`` (CPU fp32 [B, T, S], ideally pinned) holds the logits.  Two calls may be in flight, so
        the copy of the next batch overlaps the kernels of the current one.  ``x_packed`` and ``out`` must stay alive
        and untouched until the wait."""
        if x_packed.is_cuda or x_packed.dtype != torch.float32 or not x_packed.is_contiguous():
            raise FseendError("forward_host_async takes a contiguous CPU float32 tensor")
        B, T = len(ilens), int(max(ilens))
        if out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != B * T * max_nspks:
            raise FseendError("out must be a contiguous CPU float32 tensor of B * T * max_nspks elements")
        il = (C.c_int * B)(*[int(i) for i in ilens])
        ticket = C.c_longlong(-1)
        _check(self._L.fseend_fs_forward_host_async(self._h, _ptr(x_packed), il, B, max_nspks, _ptr(out),
                                                    C.byref(ticket)))
        return int(ticket.value)

    def host_wait(self, ticket: int):
        _check(self._L.fseend_fs_host_wait(self._h, int(ticket)))

    def set_option(self, key: str, value: int):
        _check(self._L.fseend_fs_set_option(self._h, key.encode(), int(value)))

    def set_profiling(self, on: bool):
        _check(self._L.fseend_fs_set_profiling(self._h, 1 if on else 0))

    def get_profile(self) -> Dict[str, Tuple[float, int]]:
        n = 64
        names = ((C.c_char * 32) * n)()
        ms = (C.c_float * n)()
        cnt = (C.c_int * n)()
        k = self._L.fseend_fs_get_profile(self._h, n, C.cast(names, C.c_void_p), ms, cnt)
        return {names[i].value.decode(): (float(ms[i]), int(cnt[i])) for i in range(k)}

    @property
    def launches_per_forward(self) -> int:
        return int(self._L.fseend_fs_launches_per_forward(self._h))

    @property
    def workspace_bytes(self) -> int:
        return int(self._L.fseend_fs_workspace_bytes(self._h))


def _pack_state_dict(state_dict):
    names, ptrs, numels, keep = [], [], [], []
    for k, v in state_dict.items():
        if not torch.is_floating_point(v):
            continue
        t = v.detach().to(device="cpu", dtype=torch.float32).contiguous()
        keep.append(t)
        names.append(k.encode())
        ptrs.append(t.data_ptr())
        numels.append(t.numel())
    n = len(names)
    return n, (C.c_char_p * n)(*names), (C.c_void_p * n)(*ptrs), (C.c_longlong * n)(*numels), keep


class LsModel:
    """Owns a native fseend_ls_model built from a reference-named LS-EEND state_dict."""

    def __init__(self, cfg: Dict[str, int], state_dict: Dict[str, torch.Tensor]):
        L = lib()
        if not torch.cuda.is_available():
            raise FseendError("fseend_b200 requires a CUDA (sm_100) device; there is no CPU fallback")
        c = LsConfig(**{k: int(cfg[k]) for k, _ in LsConfig._fields_})
        self.cfg = dict(cfg)
        n, names, ptrs, numels, keep = _pack_state_dict(state_dict)
        handle = C.c_void_p()
        _check(L.fseend_ls_create(C.byref(c), n, names, ptrs, numels, C.byref(handle)))
        self._h, self._L = handle, L

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.fseend_ls_destroy(h)

    def forward(self, x_packed: torch.Tensor, ilens: Sequence[int], max_nspks: int, want_emb: bool = False,
                want_att: bool = False):
        """Returns (logits [B,Tp,S], emb [B,Tp,D] | None, att [B,Tp,S,D] | None); Tp = max(ilens) rounded up to the
        retention chunk (the reference pads the same way)."""
        _require_cuda(x_packed)
        if x_packed.dtype != torch.float32:
            raise FseendError("x must be float32")
        B = len(ilens)
        if x_packed.shape[0] != int(sum(ilens)) or x_packed.shape[1] != self.cfg["in_size"]:
            raise FseendError("x_packed shape does not match ilens / in_size")
        Tp = int(self._L.fseend_ls_padded_len(self._h, int(max(ilens))))
        D, dev = self.cfg["n_units"], x_packed.device
        logits = torch.empty(B, Tp, max_nspks, device=dev, dtype=torch.float32)
        emb = torch.empty(B, Tp, D, device=dev, dtype=torch.float32) if want_emb else None
        att = torch.empty(B, Tp, max_nspks, D, device=dev, dtype=torch.float32) if want_att else None
        il = (C.c_int * B)(*[int(i) for i in ilens])
        _check(self._L.fseend_ls_forward(self._h, _ptr(x_packed), il, B, max_nspks, _ptr(logits), _ptr(emb), _ptr(att),
                                         _stream()))
        return logits, emb, att

    def forward_host(self, x_packed: torch.Tensor, ilens: Sequence[int], max_nspks: int,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Host-buffer entry point (H2D + forward + D2H inside the call).  x_packed: CPU fp32; returns CPU logits
        [B, Tp, S]."""
        if x_packed.is_cuda or x_packed.dtype != torch.float32 or not x_packed.is_contiguous():
            raise FseendError("forward_host takes a contiguous CPU float32 tensor")
        B = len(ilens)
        if x_packed.shape[0] != int(sum(ilens)) or x_packed.shape[1] != self.cfg["in_size"]:
            raise FseendError("x_packed shape does not match ilens / in_size")
        Tp = int(self._L.fseend_ls_padded_len(self._h, int(max(ilens))))
        logits = out if out is not None else torch.empty(B, Tp, max_nspks, dtype=torch.float32)
        il = (C.c_int * B)(*[int(i) for i in ilens])
        _check(self._L.fseend_ls_forward_host(self._h, _ptr(x_packed), il, B, max_nspks, _ptr(logits), None, None))
        return logits

    PRECISIONS = {"fp16": 0, "fp32": 1, "parity": 1}

    def set_precision(self, mode: str):
        """"fp32" / "parity" (default): fp32 activations + split-precision tcgen05 GEMMs, logits within 1e-3 of the
        reference everywhere; "fp16": fp16-operand throughput mode.  Streams created afterwards inherit it."""
        _check(self._L.fseend_ls_set_option(self._h, b"precision", self.PRECISIONS[mode]))

    @property
    def precision(self) -> str:
        return "fp32" if int(self._L.fseend_ls_get_option(self._h, b"precision")) == 1 else "fp16"

    @property
    def launches_per_forward(self) -> int:
        return int(self._L.fseend_ls_launches_per_forward(self._h))


class LsStream:
    """One-step (recurrent) LS-EEND state of B parallel recordings on top of an LsModel."""

    def __init__(self, model: "LsModel", B: int, max_nspks: int):
        self._L = lib()
        self.model, self.B, self.S = model, B, max_nspks
        h = C.c_void_p()
        _check(self._L.fseend_ls_stream_create(model._h, B, max_nspks, C.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.fseend_ls_stream_destroy(h)

    def reset(self):
        _check(self._L.fseend_ls_stream_reset(self._h))

    def step(self, x_t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """Fused frame step.  x_t: CUDA fp32 [B, in_size] or None (flush).  Returns logits [B, S] or None."""
        if x_t is not None:
            _require_cuda(x_t)
            if x_t.dtype != torch.float32 or tuple(x_t.shape) != (self.B, self.model.cfg["in_size"]):
                raise FseendError("x_t must be float32 [B, in_size]")
        out = torch.empty(self.B, self.S, device="cuda", dtype=torch.float32)
        produced = C.c_int(0)
        _check(self._L.fseend_ls_stream_step(self._h, _ptr(x_t), _ptr(out), C.byref(produced), _stream()))
        return out if produced.value else None

    def enc_step(self, x_t: torch.Tensor, t: int) -> torch.Tensor:
        _require_cuda(x_t)
        if x_t.dtype != torch.float32 or tuple(x_t.shape) != (self.B, self.model.cfg["in_size"]):
            raise FseendError("x_t must be float32 [B, in_size]")
        emb = torch.empty(self.B, self.model.cfg["n_units"], device="cuda", dtype=torch.float32)
        _check(self._L.fseend_ls_stream_enc_step(self._h, _ptr(x_t), int(t), _ptr(emb), _stream()))
        return emb

    def dec_step(self, emb: torch.Tensor, t: int) -> torch.Tensor:
        _require_cuda(emb)
        if emb.dtype != torch.float32 or tuple(emb.shape) != (self.B, self.model.cfg["n_units"]):
            raise FseendError("emb must be float32 [B, n_units]")
        att = torch.empty(self.B, self.S, self.model.cfg["n_units"], device="cuda", dtype=torch.float32)
        _check(self._L.fseend_ls_stream_dec_step(self._h, _ptr(emb), int(t), _ptr(att), _stream()))
        return att


class FsStream:
    """Device-resident streaming state of B parallel recordings on top of an FsModel."""

    def __init__(self, model: FsModel, B: int, max_nspks: int):
        self._L = lib()
        self.model = model          # keeps the native model alive
        self.B, self.S = B, max_nspks
        h = C.c_void_p()
        _check(self._L.fseend_fs_stream_create(model._h, B, max_nspks, C.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.fseend_fs_stream_destroy(h)

    def step(self, x_t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """x_t: CUDA fp32 [B, in_size] or None (flush).  Returns logits [B, S] or None while the conv window fills."""
        if x_t is not None:
            _require_cuda(x_t)
            if x_t.dtype != torch.float32 or tuple(x_t.shape) != (self.B, self.model.cfg["in_size"]):
                raise FseendError("x_t must be float32 [B, in_size]")
        out = torch.empty(self.B, self.S, device="cuda", dtype=torch.float32)
        produced = C.c_int(0)
        _check(self._L.fseend_fs_stream_step(self._h, _ptr(x_t), _ptr(out), C.byref(produced), _stream()))
        return out if produced.value else None

    @property
    def frames(self) -> int:
        return int(self._L.fseend_fs_stream_frames(self._h))


# ------------------------------------------------------------------------------ single-kernel wrappers
EPI_BIAS, EPI_LN, EPI_L2, EPI_CONVERT = 0, 1, 2, 3


def op_gemm(a: torch.Tensor, w: torch.Tensor, mode: int, *, n_seq: int = 1, taps: int = 1, tap_shift: int = 0,
            relu: bool = False, bias=None, residual=None, ln_g=None, ln_b=None, ln_eps: float = 1e-5, pe_proj=None,
            S: int = 0, seq_len=None) -> torch.Tensor:
    """a: fp16 [n_seq*rows_per_seq, K]; w: fp16 [taps*N, K]."""
    _require_cuda(a, w, bias, residual, ln_g, ln_b, pe_proj, seq_len)
    rows, K = a.shape
    N = w.shape[0] // taps
    rps = rows // n_seq
    out = torch.empty((rows, S, N) if mode == EPI_CONVERT else (rows, N), device=a.device, dtype=torch.float16)
    _check(lib().fseend_op_gemm(_ptr(a), rps, n_seq, K, _ptr(w), N, taps, tap_shift, mode, int(relu), _ptr(bias),
                                _ptr(residual), _ptr(ln_g), _ptr(ln_b), ln_eps, _ptr(pe_proj), S, _ptr(seq_len),
                                _ptr(out), _stream()))
    return out


def op_causal_attn(qkv: torch.Tensor, mask_delay: int = 0, scale: float = 0.125) -> torch.Tensor:
    """qkv: fp16 [B, T, S, 768] -> [B, T, S, 256]."""
    _require_cuda(qkv)
    B, T, S, _ = qkv.shape
    out = torch.empty(B, T, S, 256, device=qkv.device, dtype=torch.float16)
    _check(lib().fseend_op_causal_attn(_ptr(qkv), B, T, S, 4, mask_delay, scale, _ptr(out), _stream()))
    return out


def op_spk_attn(qkv: torch.Tensor, scale: float = 0.125, tensor_core: bool = False) -> torch.Tensor:
    """qkv: fp16 [frames, S, 768] -> [frames, S, 256]."""
    _require_cuda(qkv)
    F, S, _ = qkv.shape
    out = torch.empty(F, S, 256, device=qkv.device, dtype=torch.float16)
    fn = lib().fseend_op_spk_attn_tc if tensor_core else lib().fseend_op_spk_attn
    _check(fn(_ptr(qkv), F, S, scale, _ptr(out), _stream()))
    return out


def op_spk_qkv_attn(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, scale: float = 0.125) -> torch.Tensor:
    """x: fp16 [frames, S, 256]; w: fp16 [768, 256] (in_proj_weight); bias: fp32 [768] -> fp16 [frames, S, 256]."""
    _require_cuda(x, w, bias)
    F, S, _ = x.shape
    out = torch.empty(F, S, 256, device=x.device, dtype=torch.float16)
    _check(lib().fseend_op_spk_qkv_attn(_ptr(x), _ptr(w), _ptr(bias), F, S, scale, _ptr(out), _stream()))
    return out


def op_ffn(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor, ln_g: torch.Tensor,
           ln_b: torch.Tensor, *, n_seq: int = 1, ln_eps: float = 1e-5, seq_len=None, cluster: int = 2) -> torch.Tensor:
    """x: fp16 [n_seq*rows_per_seq, 256]; w1: fp16 [F, 256]; w2: fp16 [256, F]."""
    _require_cuda(x, w1, b1, w2, b2, ln_g, ln_b, seq_len)
    rows = x.shape[0]
    out = torch.empty_like(x)
    _check(lib().fseend_op_ffn(_ptr(x), rows // n_seq, n_seq, _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), w1.shape[0],
                               _ptr(ln_g), _ptr(ln_b), ln_eps, _ptr(seq_len), cluster, _ptr(out), _stream()))
    return out


def op_head(emb: torch.Tensor, att: torch.Tensor, want_f32: bool = False):
    _require_cuda(emb, att)
    F, S, _ = att.shape
    logits = torch.empty(F, S, device=emb.device, dtype=torch.float32)
    e32 = torch.empty(F, 256, device=emb.device, dtype=torch.float32) if want_f32 else None
    a32 = torch.empty(F, S, 256, device=emb.device, dtype=torch.float32) if want_f32 else None
    _check(lib().fseend_op_head(_ptr(emb), _ptr(att), F, S, _ptr(logits), _ptr(e32), _ptr(a32), _stream()))
    return logits, e32, a32


def op_prep_input(x_packed: torch.Tensor, cu_seqlens: torch.Tensor, B: int, Tmax: int, Kpad: int, scale: torch.Tensor,
                  shift: torch.Tensor) -> torch.Tensor:
    _require_cuda(x_packed, cu_seqlens, scale, shift)
    out = torch.empty(B, Tmax, Kpad, device=x_packed.device, dtype=torch.float16)
    _check(lib().fseend_op_prep_input(_ptr(x_packed), _ptr(cu_seqlens), B, Tmax, x_packed.shape[1], Kpad, _ptr(scale),
                                      _ptr(shift), _ptr(out), _stream()))
    return out


EPI_GLU, EPI_RESID = 4, 5
ACT_NONE, ACT_RELU, ACT_SWISH = 0, 1, 2


def op_gemm_ex(a: torch.Tensor, w: torch.Tensor, mode: int, *, n_seq: int = 1, act: int = 0, bias=None, residual=None,
               alpha: float = 1.0, ln_g=None, ln_b=None, ln2_g=None, ln2_b=None, ln_eps: float = 1e-5, seq_len=None):
    """Returns (out, out2 | None)."""
    _require_cuda(a, w, bias, residual, ln_g, ln_b, ln2_g, ln2_b, seq_len)
    rows, K = a.shape
    N = w.shape[0]
    n_out = N // 2 if mode == EPI_GLU else N
    out = torch.empty(rows, n_out, device=a.device, dtype=torch.float16)
    out2 = torch.empty_like(out) if ln2_g is not None else None
    _check(lib().fseend_op_gemm_ex(_ptr(a), rows // n_seq, n_seq, K, _ptr(w), N, mode, act, _ptr(bias), _ptr(residual),
                                   alpha, _ptr(ln_g), _ptr(ln_b), _ptr(ln2_g), _ptr(ln2_b), ln_eps, _ptr(seq_len),
                                   _ptr(out), _ptr(out2), _stream()))
    return out, out2


def op_retention(qkvg: torch.Tensor, chunk: int) -> torch.Tensor:
    """qkvg: fp16 [B, T, S, 1024] -> [B, T, S, 256] (swish(g) * GroupNorm(retention))."""
    _require_cuda(qkvg)
    B, T, S, _ = qkvg.shape
    nc = T // chunk
    state = torch.zeros(B * S * 4 * nc, 64, 64, device=qkvg.device, dtype=torch.float16)
    cs = torch.ones(B * S * 4 * nc, device=qkvg.device, dtype=torch.float32)
    out = torch.empty(B, T, S, 256, device=qkvg.device, dtype=torch.float16)
    _check(lib().fseend_op_retention(_ptr(qkvg), B, S, T, chunk, _ptr(state), _ptr(cs), _ptr(out), _stream()))
    return out


class P32Linear:
    """Linear layer of the parity path: fp32 weights split once into fp16 hi + lo, applied with three tcgen05 MMAs per
    product on fp32 activations.  w: fp32 [N, K] (any device; staged through the host once)."""

    def __init__(self, w: torch.Tensor):
        self._L = lib()
        wh = w.detach().to(device="cpu", dtype=torch.float32).contiguous()
        self.N, self.K = int(wh.shape[0]), int(wh.shape[1])
        if not torch.cuda.is_available():
            raise FseendError("fseend_b200 requires a CUDA (sm_100) device; there is no CPU fallback")
        h = C.c_void_p()
        _check(self._L.fseend_p32_linear_create(_ptr(wh), self.N, self.K, C.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.fseend_p32_linear_destroy(h)

    def __call__(self, a: torch.Tensor, *, bias=None, act: int = 0, alpha: float = 1.0, residual=None) -> torch.Tensor:
        _require_cuda(a, bias, residual)
        if a.dtype != torch.float32 or a.dim() != 2 or a.shape[1] != self.K:
            raise FseendError("a must be float32 [rows, K]")
        out = torch.empty(a.shape[0], self.N, device=a.device, dtype=torch.float32)
        _check(self._L.fseend_p32_linear_apply(self._h, _ptr(a), a.shape[0], _ptr(bias), int(act), float(alpha),
                                               _ptr(residual), _ptr(out), _stream()))
        return out


def op_p32_gemm(a: torch.Tensor, w: torch.Tensor, *, bias=None, act: int = 0, alpha: float = 1.0, residual=None):
    """One-shot form of P32Linear: CUDA fp32 [rows, K] x fp32 [N, K]^T -> CUDA fp32 [rows, N]."""
    lin = P32Linear(w)
    out = lin(a, bias=bias, act=act, alpha=alpha, residual=residual)
    torch.cuda.current_stream().synchronize()      # the weights are released with `lin`
    return out


def op_p32_retention(qkvg: torch.Tensor, chunk: int) -> torch.Tensor:
    """qkvg: CUDA fp32 [B, T, S, 1024] -> fp32 [B, T, S, 256]."""
    _require_cuda(qkvg)
    if qkvg.dtype != torch.float32:
        raise FseendError("qkvg must be float32")
    B, T, S, _ = qkvg.shape
    out = torch.empty(B, T, S, 256, device=qkvg.device, dtype=torch.float32)
    _check(lib().fseend_op_p32_retention(_ptr(qkvg), B, S, T, chunk, _ptr(out), _stream()))
    return out


def op_dwconv_bn_swish(u: torch.Tensor, w: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, hist=None):
    """u: fp16 [n_seq, T, 256]; w: fp32 [256, K]."""
    _require_cuda(u, w, scale, shift, hist)
    n_seq, T, _ = u.shape
    out = torch.empty_like(u)
    _check(lib().fseend_op_dwconv_bn_swish(_ptr(u), _ptr(w), _ptr(scale), _ptr(shift), n_seq, T, w.shape[1],
                                           _ptr(hist), _ptr(out), _stream()))
    return out


def op_ret_step(qkvg: torch.Tensor, state: torch.Tensor, t: int) -> torch.Tensor:
    """qkvg: fp16 [n_seq, 1024]; state: fp32 [n_seq, 4, 64, 64] (updated in place)."""
    _require_cuda(qkvg, state)
    out = torch.empty(qkvg.shape[0], 256, device=qkvg.device, dtype=torch.float16)
    _check(lib().fseend_op_ret_step(_ptr(qkvg), _ptr(state), qkvg.shape[0], t, _ptr(out), _stream()))
    return out


def op_embloss(emb: torch.Tensor, labels: torch.Tensor, seq_len: Optional[torch.Tensor] = None,
               divisor: Optional[float] = None) -> torch.Tensor:
    """Embedding-consistency loss.  emb: fp32 [B, T, 256]; labels: fp32 [B, T, S] zero padded; seq_len: int32 [B] on the
    device (LS-EEND masking) or None (FS-EEND: every padded row counts).  Returns a 0-dim fp32 CUDA tensor."""
    _require_cuda(emb, labels, seq_len)
    if emb.dtype != torch.float32 or labels.dtype != torch.float32:
        raise FseendError("emb and labels must be float32")
    B, T, D = emb.shape
    if D != 256 or tuple(labels.shape[:2]) != (B, T):
        raise FseendError("emb must be [B, T, 256] and labels [B, T, S]")
    S = labels.shape[2]
    if divisor is None:
        divisor = float(B) * T * T
    L = lib()
    ws = torch.empty(int(L.fseend_op_embloss_workspace_bytes(B, T)) // 4, device=emb.device, dtype=torch.float32)
    loss = torch.empty((), device=emb.device, dtype=torch.float32)
    _check(L.fseend_op_embloss(_ptr(emb), _ptr(labels), _ptr(seq_len), B, T, S, float(divisor), _ptr(ws), _ptr(loss),
                               _stream()))
    return loss


def op_decide_median(pred: torch.Tensor, threshold: float = 0.5, median: int = 11) -> torch.Tensor:
    """pred: CUDA fp32 [T, C] posteriors -> uint8 [T, C] decisions = medfilt(pred > threshold, (median, 1))."""
    _require_cuda(pred)
    if pred.dtype != torch.float32 or pred.dim() != 2:
        raise FseendError("pred must be float32 [T, C]")
    T, Cn = pred.shape
    out = torch.empty(T, Cn, device=pred.device, dtype=torch.uint8)
    _check(lib().fseend_op_decide_median(_ptr(pred), T, Cn, float(threshold), int(median), _ptr(out), _stream()))
    return out


def op_label_prepare(labels: torch.Tensor):
    """labels: CUDA fp32 [B, T, n_spk] 0/1 -> (labels_out [B, T, n_spk + 2], perm int32 [B, n_spk])."""
    _require_cuda(labels)
    if labels.dtype != torch.float32 or labels.dim() != 3:
        raise FseendError("labels must be float32 [B, T, n_spk]")
    B, T, Cn = labels.shape
    out = torch.empty(B, T, Cn + 2, device=labels.device, dtype=torch.float32)
    perm = torch.empty(B, Cn, device=labels.device, dtype=torch.int32)
    _check(lib().fseend_op_label_prepare(_ptr(labels), B, T, Cn, _ptr(perm), _ptr(out), _stream()))
    return out, perm


def op_bce_loss(logits: torch.Tensor, target: torch.Tensor, lens: torch.Tensor, n_cls: torch.Tensor,
                label_delay: int = 0) -> torch.Tensor:
    """logits [B, T, Cy], target [B, T, Ct] CUDA fp32 (padded); lens, n_cls int32 [B] on the device."""
    _require_cuda(logits, target, lens, n_cls)
    if logits.dtype != torch.float32 or target.dtype != torch.float32:
        raise FseendError("logits and target must be float32")
    if lens.dtype != torch.int32 or n_cls.dtype != torch.int32:
        raise FseendError("lens and n_cls must be int32")
    B, T, Cy = logits.shape
    if tuple(target.shape[:2]) != (B, T):
        raise FseendError("target must be [B, T, C]")
    L = lib()
    ws = torch.empty(max(1, int(L.fseend_op_bce_loss_workspace_bytes(B, T)) // 4), device=logits.device, dtype=torch.float32)
    loss = torch.empty((), device=logits.device, dtype=torch.float32)
    _check(L.fseend_op_bce_loss(_ptr(logits), Cy, _ptr(target), target.shape[2], B, T, _ptr(lens), _ptr(n_cls),
                                int(label_delay), _ptr(ws), _ptr(loss), _stream()))
    return loss


def op_pit_costs(logits: torch.Tensor, labels: torch.Tensor, lens: torch.Tensor, label_delay: int = 0,
                 pad_term: bool = False) -> torch.Tensor:
    """logits, labels: CUDA fp32 [B, T, C] (zero padded), lens int32 [B] on the device -> fp64 [B, C, C] pair costs."""
    _require_cuda(logits, labels, lens)
    if logits.dtype != torch.float32 or labels.dtype != torch.float32 or lens.dtype != torch.int32:
        raise FseendError("logits / labels must be float32 and lens int32")
    if logits.shape != labels.shape or logits.dim() != 3:
        raise FseendError("logits and labels must both be [B, T, C]")
    B, T, Cn = logits.shape
    cost = torch.empty(B, Cn, Cn, device=logits.device, dtype=torch.float64)
    _check(lib().fseend_op_pit_costs(_ptr(logits), _ptr(labels), B, T, Cn, _ptr(lens), int(label_delay),
                                     1 if pad_term else 0, _ptr(cost), _stream()))
    return cost


def op_splice_subsample(feat: torch.Tensor, context_size: int = 7, subsampling: int = 10) -> torch.Tensor:
    """feat: CUDA fp32 [T, F] -> [ceil(T / subsampling), (2 * context_size + 1) * F] (splice then subsample)."""
    _require_cuda(feat)
    if feat.dtype != torch.float32 or feat.dim() != 2:
        raise FseendError("feat must be float32 [T, F]")
    T, F = feat.shape
    out = torch.empty((T + subsampling - 1) // subsampling, (2 * context_size + 1) * F, device=feat.device,
                      dtype=torch.float32)
    _check(lib().fseend_op_splice_subsample(_ptr(feat), T, F, int(context_size), int(subsampling), _ptr(out), _stream()))
    return out
