"""ctypes binding of libb200vqa.so (include/b200vqa.h).

This is the only place Python touches the C ABI.  There is no CPU or PyTorch fallback: if the shared
library is missing or the device is not an sm_100 GPU every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200vqa.so")

MODEL_IQAP = 0
MODEL_FA = 1

_f32p = C.POINTER(C.c_float)
_vp = C.c_void_p


class NativeError(RuntimeError):
    """A libb200vqa call returned a negative status."""


class MhaWeights(C.Structure):
    _fields_ = [("in_proj_weight", _vp), ("in_proj_bias", _vp), ("out_proj_weight", _vp), ("out_proj_bias", _vp)]


class EncoderLayerWeights(C.Structure):
    _fields_ = [("self_attn", MhaWeights),
                ("linear1_weight", _vp), ("linear1_bias", _vp), ("linear2_weight", _vp), ("linear2_bias", _vp),
                ("norm1_weight", _vp), ("norm1_bias", _vp), ("norm2_weight", _vp), ("norm2_bias", _vp)]


class DecoderLayerWeights(C.Structure):
    _fields_ = [("self_attn", MhaWeights), ("multihead_attn", MhaWeights),
                ("linear1_weight", _vp), ("linear1_bias", _vp), ("linear2_weight", _vp), ("linear2_bias", _vp),
                ("norm1_weight", _vp), ("norm1_bias", _vp), ("norm2_weight", _vp), ("norm2_bias", _vp),
                ("norm3_weight", _vp), ("norm3_bias", _vp)]


class ModelDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("d_model", C.c_int32), ("img_feat_dim", C.c_int32), ("n_img_tokens", C.c_int32),
                ("nhead", C.c_int32), ("n_enc_layers", C.c_int32), ("n_dec_layers", C.c_int32), ("dim_ff", C.c_int32),
                ("enc_vocab", C.c_int32), ("dec_vocab", C.c_int32), ("max_q_len", C.c_int32),
                ("pe_enc_len", C.c_int32), ("pe_dec_len", C.c_int32), ("answer_hidden", C.c_int32),
                ("num_classes", C.c_int32), ("layer_norm_eps", C.c_float), ("answer_pool_rows", C.c_int32),
                ("image_proj_weight", _vp), ("image_proj_bias", _vp), ("cls_token", _vp),
                ("enc_embedding", _vp), ("dec_embedding", _vp), ("pe_enc", _vp), ("pe_dec", _vp),
                ("enc_layers", C.POINTER(EncoderLayerWeights)), ("dec_layers", C.POINTER(DecoderLayerWeights)),
                ("enc_final_norm_weight", _vp), ("enc_final_norm_bias", _vp),
                ("dec_final_norm_weight", _vp), ("dec_final_norm_bias", _vp),
                ("head_weight", _vp), ("head_bias", _vp),
                ("answer_w0", _vp), ("answer_b0", _vp), ("answer_w1", _vp), ("answer_b1", _vp)]


class LstmDesc(C.Structure):
    _fields_ = [("vocab", C.c_int32), ("embedding_dim", C.c_int32), ("hidden_dim", C.c_int32), ("prog_vocab", C.c_int32),
                ("embedding", _vp), ("enc_w_ih", _vp), ("enc_w_hh", _vp), ("enc_b_ih", _vp), ("enc_b_hh", _vp),
                ("dec_w_ih", _vp), ("dec_w_hh", _vp), ("dec_b_ih", _vp), ("dec_b_hh", _vp), ("fc_w", _vp), ("fc_b", _vp)]


class DbgGemmArgs(C.Structure):
    _fields_ = [("epilogue", C.c_int32), ("tf32", C.c_int32), ("block_n", C.c_int32),
                ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
                ("A", _vp), ("W", _vp), ("bias", _vp), ("out", _vp), ("ldc", C.c_int32),
                ("residual", _vp), ("gamma", _vp), ("beta", _vp), ("out_f32", _vp),
                ("rows_in", C.c_int32), ("rows_out", C.c_int32), ("row_off", C.c_int32), ("pe_off", C.c_int32),
                ("pe", _vp), ("clk", _vp)]


# name -> (restype, argtypes); the not-gpu test-suite checks every symbol of include/b200vqa.h is here and exported
SIGNATURES = {
    "b200vqa_create": (C.c_int, [C.POINTER(ModelDesc), C.c_int, C.POINTER(_vp)]),
    "b200vqa_destroy": (None, [_vp]),
    "b200vqa_refresh_weights": (C.c_int, [_vp, C.POINTER(ModelDesc)]),
    "b200vqa_workspace_bytes": (C.c_size_t, [_vp, C.c_int]),
    "b200vqa_last_error": (C.c_char_p, []),
    "b200vqa_version": (C.c_char_p, []),
    "b200vqa_launch_count": (C.c_uint64, [_vp]),
    "b200vqa_set_start_token": (C.c_int, [_vp, C.c_int]),
    "b200vqa_profile_num_tags": (C.c_int, []),
    "b200vqa_profile_tag_name": (C.c_char_p, [C.c_int]),
    "b200vqa_profile_begin": (C.c_int, [_vp]),
    "b200vqa_profile_end": (C.c_int, [_vp, _vp, _vp]),
    "b200vqa_profile_delay": (C.c_int, [_vp, C.c_double, _vp]),
    "b200vqa_iqap_forward": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200vqa_iqap_forward_indexed": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "b200vqa_iqap_forward_host_indexed": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int,
                                                    _vp]),
    "b200vqa_iqap_forward_f16": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200vqa_iqap_forward_host_f16": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp]),
    "b200vqa_iqap_tally": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "b200vqa_iqap_decode": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "b200vqa_iqap_forward_host": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp]),
    "b200vqa_iqap_forward_host_async": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp]),
    "b200vqa_fa_project_images": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "b200vqa_fa_step": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "b200vqa_fa_forward": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp]),
    "b200vqa_fa_run_chain": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp,
                                       _vp, _vp]),
    "b200vqa_fa_run_chain_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int,
                                            _vp]),
    "b200vqa_fa_run_chain_host_async": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp,
                                                  C.c_int, _vp]),
    "b200vqa_fa_run_chain_indexed": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int,
                                               C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "b200vqa_lstm_create": (C.c_int, [C.POINTER(LstmDesc), C.c_int, C.POINTER(_vp)]),
    "b200vqa_lstm_destroy": (None, [_vp]),
    "b200vqa_lstm_launch_count": (C.c_uint64, [_vp]),
    "b200vqa_lstm_generate": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "b200vqa_programs_to_chain": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "b200vqa_dbg_gemm": (C.c_int, [C.POINTER(DbgGemmArgs), _vp]),
    "b200vqa_dbg_gemm_check": (C.c_int, [C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "b200vqa_dbg_enc_attention": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "b200vqa_dbg_workspace": (C.c_int, [_vp, C.c_int, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "b200vqa_host_f32_to_f16": (C.c_int, [_vp, _vp, C.c_longlong, C.c_int]),
    "b200vqa_host_f32_to_bf16": (C.c_int, [_vp, _vp, C.c_longlong, C.c_int]),
    "b200vqa_set_host_upload": (C.c_int, [_vp, C.c_int]),
    "b200vqa_dbg_mem_attn": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
}

_lib = None
_lib_lock = threading.Lock()


def lib() -> C.CDLL:
    """Loads libb200vqa.so (built in-tree by `__graft_entry__.build()` / `make -C .../csrc`)."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise NativeError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(there is no CPU / PyTorch fallback for the executor path)")
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)
                fn.restype = res
                fn.argtypes = args
            _lib = handle
    return _lib


def last_error() -> str:
    msg = lib().b200vqa_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise NativeError(f"{what} failed with status {rc}: {last_error()}")


def ptr(t):
    """Raw device (or host) pointer of a tensor, NULL for None."""
    return _vp(0) if t is None else _vp(t.data_ptr())


def stream_ptr(device=None):
    return _vp(torch.cuda.current_stream(device).cuda_stream)


def _dev_f32(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise NativeError(f"{what} lives on {t.device}: the executor path runs on an sm_100 GPU only (no CPU fallback); "
                          "move the module with .cuda()")
    t = t.detach()
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(torch.float32).contiguous()
    return t


class Handle:
    """Owns one b200vqa_handle: packed weights + activation workspace on one GPU."""

    def __init__(self, desc_builder, device: torch.device):
        self._lib = lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise NativeError(f"device {self.device}: libb200vqa needs an sm_100 GPU (no CPU fallback)")
        self._h = _vp(0)
        desc, keep = desc_builder()
        self._keep = keep
        out = _vp(0)
        with torch.cuda.device(self.device):
            torch.cuda.current_stream().synchronize()  # parameter writes must be visible to the packer
            check(self._lib.b200vqa_create(C.byref(desc), self.device.index or 0, C.byref(out)), "b200vqa_create")
        self._h = out

    @property
    def raw(self):
        return self._h

    def refresh(self, desc_builder):
        desc, keep = desc_builder()
        with torch.cuda.device(self.device):
            torch.cuda.current_stream().synchronize()
            check(self._lib.b200vqa_refresh_weights(self._h, C.byref(desc)), "b200vqa_refresh_weights")
        self._keep = keep

    def launch_count(self) -> int:
        return int(self._lib.b200vqa_launch_count(self._h))

    def set_host_upload(self, fp16: bool) -> None:
        check(self._lib.b200vqa_set_host_upload(self._h, 1 if fp16 else 0), "b200vqa_set_host_upload")

    def set_start_token(self, token: int) -> None:
        check(self._lib.b200vqa_set_start_token(self._h, int(token)), "b200vqa_set_start_token")

    def workspace_bytes(self, B: int) -> int:
        return int(self._lib.b200vqa_workspace_bytes(self._h, int(B)))

    def profile_begin(self):
        check(self._lib.b200vqa_profile_begin(self._h), "b200vqa_profile_begin")

    def profile_delay(self, ms: float):
        check(self._lib.b200vqa_profile_delay(self._h, float(ms), stream_ptr(self.device)), "b200vqa_profile_delay")

    def profile_end(self) -> dict:
        """{kernel class: (total ms, launches)} since profile_begin (synchronises the device)."""
        n = self._lib.b200vqa_profile_num_tags()
        ms = (C.c_float * n)()
        cnt = (C.c_int32 * n)()
        check(self._lib.b200vqa_profile_end(self._h, ms, cnt), "b200vqa_profile_end")
        return {self._lib.b200vqa_profile_tag_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i]}

    def dbg_workspace(self, which: int, dtype: torch.dtype) -> torch.Tensor:
        """Copy of one decode scratch buffer (test hook, see b200vqa_dbg_workspace)."""
        n = C.c_size_t(0)
        check(self._lib.b200vqa_dbg_workspace(self._h, int(which), _vp(0), 0, C.byref(n)), "b200vqa_dbg_workspace")
        out = torch.empty(n.value // torch.empty((), dtype=dtype).element_size(), dtype=dtype, device=self.device)
        check(self._lib.b200vqa_dbg_workspace(self._h, int(which), _vp(out.data_ptr()), n.value, C.byref(n)),
              "b200vqa_dbg_workspace")
        return out

    def close(self):
        if self._h:
            self._lib.b200vqa_destroy(self._h)
            self._h = _vp(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def mha_weights(mha, keep) -> MhaWeights:
    w = MhaWeights()
    for name in ("in_proj_weight", "in_proj_bias"):
        t = _dev_f32(getattr(mha, name), name)
        keep.append(t)
        setattr(w, name, t.data_ptr())
    t = _dev_f32(mha.out_proj.weight, "out_proj.weight"); keep.append(t); w.out_proj_weight = t.data_ptr()
    t = _dev_f32(mha.out_proj.bias, "out_proj.bias"); keep.append(t); w.out_proj_bias = t.data_ptr()
    return w


def _set(struct, keep, **tensors):
    for name, t in tensors.items():
        if t is None:
            setattr(struct, name, None)
            continue
        t = _dev_f32(t, name)
        keep.append(t)
        setattr(struct, name, t.data_ptr())


def encoder_layer_weights(layer, keep) -> EncoderLayerWeights:
    w = EncoderLayerWeights()
    w.self_attn = mha_weights(layer.self_attn, keep)
    _set(w, keep, linear1_weight=layer.linear1.weight, linear1_bias=layer.linear1.bias,
         linear2_weight=layer.linear2.weight, linear2_bias=layer.linear2.bias,
         norm1_weight=layer.norm1.weight, norm1_bias=layer.norm1.bias,
         norm2_weight=layer.norm2.weight, norm2_bias=layer.norm2.bias)
    return w


def decoder_layer_weights(layer, keep) -> DecoderLayerWeights:
    w = DecoderLayerWeights()
    w.self_attn = mha_weights(layer.self_attn, keep)
    w.multihead_attn = mha_weights(layer.multihead_attn, keep)
    _set(w, keep, linear1_weight=layer.linear1.weight, linear1_bias=layer.linear1.bias,
         linear2_weight=layer.linear2.weight, linear2_bias=layer.linear2.bias,
         norm1_weight=layer.norm1.weight, norm1_bias=layer.norm1.bias,
         norm2_weight=layer.norm2.weight, norm2_bias=layer.norm2.bias,
         norm3_weight=layer.norm3.weight, norm3_bias=layer.norm3.bias)
    return w


def check_layer_contract(layer, what: str) -> None:
    """The kernels implement torch's defaults the reference relies on: post-norm, ReLU."""
    import torch.nn.functional as F
    if getattr(layer, "norm_first", False):
        raise NativeError(f"{what}: norm_first=True is not implemented (the reference uses post-norm)")
    act = getattr(layer, "activation", F.relu)
    if act is not F.relu and not isinstance(act, torch.nn.ReLU):
        raise NativeError(f"{what}: only the ReLU activation of the reference is implemented")


def weights_version(module: torch.nn.Module):
    """Changes whenever a parameter/buffer is modified in place or replaced (load_state_dict, .to(), .cuda())."""
    return tuple((id(t), t._version, t.data_ptr()) for t in list(module.parameters()) + list(module.buffers()))


class HandlePool:
    """Per-module pool of native handles ("slots").  Slot 0 serves the plain reference-style calls on the caller's
    stream.  Further slots each own a handle (packed weights, workspace, CUDA graphs) and a stream, so independent
    batches can be in flight at once: the decode phase is a latency-bound chain of small kernels, and a second
    batch's kernels fill the bubbles (`submit` / `drain` on the modules)."""

    def __init__(self, module, desc_builder):
        self._module = module
        self._builder = desc_builder
        self._handles = {}
        self._versions = {}
        self._streams = {}
        self._workers = {}   # slot -> single-thread executor (background host-buffer submissions)
        self.futures = []    # pending background calls

    def __getstate__(self):
        # native handles, streams, worker threads and packed weights never travel through pickle / deepcopy: they are
        # rebuilt lazily
        return {"_module": self._module, "_builder": self._builder, "_handles": {}, "_versions": {}, "_streams": {},
                "_workers": {}, "futures": []}

    def __setstate__(self, state):
        self.__dict__.update(state)

    def get(self, slot: int = 0) -> Handle:
        module = self._module
        version = weights_version(module)
        dev = next(module.parameters()).device
        h = self._handles.get(slot)
        if h is not None and h.device != dev:
            h.close()
            h = None
        if h is None:
            h = Handle(self._builder, dev)
            self._handles[slot] = h
        elif self._versions.get(slot) != version:
            h.refresh(self._builder)
        self._versions[slot] = version
        return h

    def stream(self, slot: int) -> "torch.cuda.Stream":
        dev = next(self._module.parameters()).device
        st = self._streams.get(slot)
        if st is None or st.device != dev:
            st = torch.cuda.Stream(device=dev)
            self._streams[slot] = st
        return st

    def run_in_background(self, slot: int, fn) -> None:
        """Runs fn() on the slot's own host thread (one per slot: calls on a handle stay in submission order; ctypes
        releases the GIL inside the library).  `wait_background()` re-raises its exception."""
        import concurrent.futures
        ex = self._workers.get(slot)
        if ex is None:
            ex = self._workers[slot] = concurrent.futures.ThreadPoolExecutor(max_workers=1)
        self.futures.append(ex.submit(fn))

    def wait_background(self) -> None:
        futures, self.futures = self.futures, []
        for f in futures:
            f.result()

    def drain(self):
        """Makes the caller's current stream wait for every slot stream (results of all submits become usable)."""
        for st in self._streams.values():
            torch.cuda.current_stream(st.device).wait_stream(st)

    def launch_count(self) -> int:
        return sum(h.launch_count() for h in self._handles.values())
