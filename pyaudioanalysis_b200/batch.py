"""Batched device API: [B, N] clips resident in HBM -> feature tensors resident in HBM.

torch supplies device memory and the current stream; every operation is a C-ABI call into
libb200aa.so (hand-written sm_100a kernels).  Nothing here computes on the CPU.
"""
import ctypes

import torch

from . import _lib
from ._lib import lib, check, get_plan, DTYPE_I16, DTYPE_F32

NORM_BYTES = 32


class DeviceBuffer:
    """Raw float32 device memory [shape] given by address -- e.g. a window of another GPU's HBM mapped over NVLink
    (dist.PeerGather).  Accepted as ``out=`` of feature_extraction_batch; the caller guarantees size and lifetime."""

    def __init__(self, ptr, shape):
        self.ptr, self.shape, self.is_cuda = int(ptr), tuple(int(s) for s in shape), True

    def data_ptr(self):
        return self.ptr


def _require_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("%s must be a CUDA tensor (there is no CPU fallback)" % name)


def _dtype_code(t):
    if t.dtype == torch.int16:
        return DTYPE_I16
    if t.dtype == torch.float32:
        return DTYPE_F32
    raise TypeError("clips must be int16 or float32, got %s" % t.dtype)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _prep(signals, lengths):
    _require_cuda(signals, "signals")
    if signals.dim() == 1:
        signals = signals.unsqueeze(0)
    if signals.dim() != 2 or signals.stride(1) != 1:
        raise ValueError("signals must be [B, N] with contiguous samples")
    B, N = signals.shape
    stride = signals.stride(0) if B > 1 else N
    if stride < N:
        raise ValueError("overlapping clips are not supported")
    len_ptr = None
    if lengths is not None:
        _require_cuda(lengths, "lengths")
        lengths = lengths.to(torch.int64).contiguous()
        if lengths.numel() != B:
            raise ValueError("lengths must have one entry per clip")
        len_ptr = ctypes.c_void_p(lengths.data_ptr())
    return signals, B, N, stride, lengths, len_ptr


def clip_stats(signals, lengths=None, out=None):
    """Per-clip normalisation records (kernel 0).  Returns a uint8 CUDA tensor [B, 32]."""
    signals, B, N, stride, lengths, len_ptr = _prep(signals, lengths)
    with torch.cuda.device(signals.device):
        norm = out if out is not None else torch.empty((B, NORM_BYTES), dtype=torch.uint8, device=signals.device)
        check(lib().b200aa_clip_stats(ctypes.c_void_p(signals.data_ptr()), _dtype_code(signals), B, N, stride,
                                      len_ptr, ctypes.c_void_p(norm.data_ptr()), _stream()))
    return norm


def feature_extraction_batch(signals, sampling_rate, window, step, deltas=True, out=None, lengths=None, norm=None,
                             plan=None):
    """Short-term features of B clips: CUDA [B, N] int16/float32 -> CUDA float32 [B, 68|34, T].

    Same semantics per clip as ShortTermFeatures.feature_extraction (reference
    ShortTermFeatures.py:543-685); ``lengths`` (int64 CUDA [B]) allows ragged clips: columns
    beyond a clip's own frame count are left as they are in ``out`` (zeros if allocated here).
    """
    window, step = int(window), int(step)
    signals, B, N, stride, lengths, len_ptr = _prep(signals, lengths)
    F = 68 if deltas else 34
    with torch.cuda.device(signals.device):
        plan = plan or get_plan(sampling_rate, window, step, signals.device.index)
        T = lib().b200aa_num_frames(N, window, step)
        if T <= 0:
            check(_lib.ERR_TOO_SHORT)
        if out is None:
            alloc = torch.zeros if lengths is not None else torch.empty
            out = alloc((B, F, T), dtype=torch.float32, device=signals.device)
        elif isinstance(out, DeviceBuffer):
            if out.shape[0] != B or out.shape[1] != F or out.shape[2] < T:
                raise ValueError("out must be float32 [B, %d, >=%d]" % (F, T))
        else:
            _require_cuda(out, "out")
            if out.dtype != torch.float32 or out.shape[0] != B or out.shape[1] != F or out.shape[2] < T or not out.is_contiguous():
                raise ValueError("out must be contiguous float32 [B, %d, >=%d]" % (F, T))
        if norm is None:
            norm = clip_stats(signals, lengths)
        check(lib().b200aa_st_features(plan.handle, ctypes.c_void_p(signals.data_ptr()), _dtype_code(signals), B, N, stride,
                                       len_ptr, ctypes.c_void_p(norm.data_ptr()), 1 if deltas else 0,
                                       ctypes.c_void_p(out.data_ptr()), out.shape[2], _stream()))
    return out


def mid_pool_batch(st, ratio, step_ratio, n_frames=None):
    """Mean / population-std pooling (kernel 2): CUDA float32 [B, F, T] -> [B, 2F, M]."""
    _require_cuda(st, "st")
    if st.dim() != 3 or st.dtype != torch.float32 or not st.is_contiguous():
        raise ValueError("st must be contiguous float32 [B, F, T]")
    B, F, Tst = st.shape
    T = Tst if n_frames is None else int(n_frames)
    M = lib().b200aa_mid_windows(T, int(step_ratio))
    with torch.cuda.device(st.device):
        mid = torch.empty((B, 2 * F, M), dtype=torch.float32, device=st.device)
        check(lib().b200aa_mid_pool(ctypes.c_void_p(st.data_ptr()), B, F, T, Tst, int(ratio), int(step_ratio),
                                    ctypes.c_void_p(mid.data_ptr()), _stream()))
    return mid


def long_term_mean_batch(mid):
    """Mean over the mid-term windows: CUDA float32 [B, rows, M] -> [B, rows] (MidTermFeatures.py:200-201)."""
    _require_cuda(mid, "mid")
    if mid.dim() != 3 or mid.dtype != torch.float32 or not mid.is_contiguous():
        raise ValueError("mid must be contiguous float32 [B, rows, M]")
    B, rows, M = mid.shape
    with torch.cuda.device(mid.device):
        out = torch.empty((B, rows), dtype=torch.float32, device=mid.device)
        check(lib().b200aa_long_term_mean(ctypes.c_void_p(mid.data_ptr()), B, rows, M, ctypes.c_void_p(out.data_ptr()), _stream()))
    return out


def mid_ratios(mid_window, mid_step, short_window, short_step):
    """MidTermFeatures.py:100-102 (Python round(): half to even).  window/step truncation happens
    inside feature_extraction only (ShortTermFeatures.py:563-564); the ratios use the raw arguments."""
    ratio = round((mid_window - (short_window - short_step)) / short_step)
    stepr = int(round(mid_step / short_step))
    return int(ratio), stepr


def mid_feature_extraction_batch(signals, sampling_rate, mid_window, mid_step, short_window, short_step):
    """Batched MidTermFeatures.mid_feature_extraction: returns (mid [B,136,M], st [B,68,T]) on the GPU."""
    st = feature_extraction_batch(signals, sampling_rate, short_window, short_step, deltas=True)
    ratio, stepr = mid_ratios(mid_window, mid_step, short_window, short_step)
    if ratio < 1 or stepr < 1:
        raise ValueError("mid-term window / step shorter than one short-term step")
    return mid_pool_batch(st, ratio, stepr), st


def spectrogram_batch(signals, sampling_rate, window, step, plan=None, norm=None, out=None):
    """CUDA [B, N] -> CUDA float32 [B, R, K] (ShortTermFeatures.py:389-452 rows, per clip).  ``norm``: records of a previous
    ``clip_stats`` call on the same clips; ``out``: a contiguous float32 [B, R, K] tensor to write into."""
    window, step = int(window), int(step)
    signals, B, N, stride, _, _ = _prep(signals, None)
    with torch.cuda.device(signals.device):
        plan = plan or get_plan(sampling_rate, window, step, signals.device.index)
        R = lib().b200aa_spectrogram_rows(N, window, step)
        if R <= 0:
            check(_lib.ERR_TOO_SHORT)
        if out is None:
            out = torch.empty((B, R, window // 2), dtype=torch.float32, device=signals.device)
        elif tuple(out.shape) != (B, R, window // 2) or out.dtype != torch.float32 or not out.is_contiguous() or not out.is_cuda:
            raise ValueError("out must be a contiguous float32 CUDA tensor [B, %d, %d]" % (R, window // 2))
        if norm is None:
            norm = clip_stats(signals)
        check(lib().b200aa_spectrogram(plan.handle, ctypes.c_void_p(signals.data_ptr()), _dtype_code(signals), B, N, stride,
                                       ctypes.c_void_p(norm.data_ptr()), ctypes.c_void_p(out.data_ptr()), _stream()))
    return out


def chromagram_batch(signals, sampling_rate, window, step, plan=None, norm=None):
    """CUDA [B, N] -> CUDA float32 [B, R, 12] (ShortTermFeatures.py:324-386 rows, per clip)."""
    window, step = int(window), int(step)
    signals, B, N, stride, _, _ = _prep(signals, None)
    with torch.cuda.device(signals.device):
        plan = plan or get_plan(sampling_rate, window, step, signals.device.index)
        R = lib().b200aa_chromagram_rows(N, window, step)
        if R <= 0 or N - step - window < 0:
            check(_lib.ERR_TOO_SHORT)
        out = torch.empty((B, R, 12), dtype=torch.float32, device=signals.device)
        if norm is None:
            norm = clip_stats(signals)
        check(lib().b200aa_chromagram(plan.handle, ctypes.c_void_p(signals.data_ptr()), _dtype_code(signals), B, N, stride,
                                      ctypes.c_void_p(norm.data_ptr()), ctypes.c_void_p(out.data_ptr()), _stream()))
    return out
