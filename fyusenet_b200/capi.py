"""ctypes binding of the C ABI declared in include/fyusenet_b200.h.

This is the Python-side stub a maintainer would use to drive the backend; it is
what tests/ and bench.py call.  There is no fallback of any kind: if libfyusenet_b200.so is missing
the import fails loudly, and without a CUDA device `Context()` raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "libfyusenet_b200.so"

# flags / enums (include/fyusenet_b200.h)
FLAG_RESIDUAL_INPUT, FLAG_RELU_ON_RESIDUAL, FLAG_BATCHNORM_ON_RESIDUAL = 1, 2, 4
FLAG_POST_BATCHNORM, FLAG_DEEP, FLAG_PRE_RELU, FLAG_PRE_CLIP = 8, 16, 64, 128
QUIRK_FRAC3_ASYM, QUIRK_FRAC_ACT_FIRST, QUIRK_MAXPOOL3_COL, QUIRK_DW_BN_OFFSET, QUIRK_TRANS2X2_NEXT, QUIRKS_REFERENCE = 1, 2, 4, 8, 16, 31
ORDER_SHALLOW, ORDER_DEEP = 0, 1
F16, F32 = 0, 1
BACKEND_AUTO, BACKEND_DIRECT, BACKEND_TC = 0, 1, 2


class FynError(RuntimeError):
    """Non-zero status from the C ABI (the C++ host wrapper throws FynException for the same)."""


class TensorDesc(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("channels", C.c_int), ("padding", C.c_int),
                ("order", C.c_int), ("dtype", C.c_int), ("batch", C.c_int), ("packing", C.c_int)]


class TensorGeom(C.Structure):
    _fields_ = [("tex_width", C.c_int), ("tex_height", C.c_int), ("planes", C.c_int), ("tiles_x", C.c_int),
                ("tiles_y", C.c_int), ("packing", C.c_int), ("elem_size", C.c_size_t),
                ("plane_elems", C.c_size_t), ("image_elems", C.c_size_t), ("bytes", C.c_size_t)]


class DeviceInfo(C.Structure):
    _fields_ = [("device", C.c_int), ("sm_count", C.c_int), ("cc_major", C.c_int), ("cc_minor", C.c_int),
                ("total_mem", C.c_size_t), ("smem_per_block_optin", C.c_size_t), ("name", C.c_char * 64)]


class ConvPlanInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("mode", "n", "steps", "window_rows", "row_advance", "phases_x", "phases_y", "ring_slots",
                                       "mirror_slots", "slot_bytes", "staged_rows", "stage_bytes", "row_items", "loader_groups",
                                       "epilogue_warps", "bias_folded")] + [("weight_image_bytes", C.c_size_t), ("shared_bytes", C.c_size_t)]


class ConvDesc(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("in_channels", C.c_int), ("out_channels", C.c_int),
                ("kernel", C.c_int), ("downsample", C.c_int), ("dilation", C.c_int),
                ("in_padding", C.c_int), ("out_padding", C.c_int), ("res_padding", C.c_int),
                ("flags", C.c_uint), ("leaky", C.c_float), ("clip_lo", C.c_float), ("clip_hi", C.c_float),
                ("source_step", C.c_float), ("fractional", C.c_int), ("quirks", C.c_int), ("backend", C.c_int)]


class PoolDesc(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("channels", C.c_int), ("pool_x", C.c_int),
                ("pool_y", C.c_int), ("downsample", C.c_int), ("in_padding", C.c_int), ("out_padding", C.c_int),
                ("is_max", C.c_int), ("global_", C.c_int), ("flags", C.c_uint), ("leaky", C.c_float),
                ("clip_lo", C.c_float), ("clip_hi", C.c_float), ("quirks", C.c_int)]


class BnDesc(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("channels", C.c_int), ("in_padding", C.c_int),
                ("out_padding", C.c_int), ("flags", C.c_uint), ("leaky", C.c_float), ("clip_lo", C.c_float),
                ("clip_hi", C.c_float)]


UnaryDesc = BnDesc  # identical field layout (fyn_unary_desc)


class ScaleDesc(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("channels", C.c_int), ("in_padding", C.c_int),
                ("out_padding", C.c_int), ("upsample_x", C.c_int), ("upsample_y", C.c_int), ("downsample_x", C.c_int),
                ("downsample_y", C.c_int), ("linear", C.c_int), ("flags", C.c_uint), ("leaky", C.c_float),
                ("clip_lo", C.c_float), ("clip_hi", C.c_float)]


class ArithDesc(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("channels", C.c_int), ("in_padding", C.c_int),
                ("out_padding", C.c_int), ("op", C.c_int), ("singleton", C.c_int), ("operand", C.c_float),
                ("flags", C.c_uint), ("leaky", C.c_float), ("clip_lo", C.c_float), ("clip_hi", C.c_float)]


CONCAT_MAX_INPUTS = 8


class ConcatDesc(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("num_inputs", C.c_int), ("channels", C.c_int * CONCAT_MAX_INPUTS),
                ("in_padding", C.c_int), ("out_padding", C.c_int), ("flags", C.c_uint), ("leaky", C.c_float),
                ("clip_lo", C.c_float), ("clip_hi", C.c_float)]


ARITH_ADD, ARITH_SUB, ARITH_MUL, ARITH_DIV = 0, 1, 2, 3


class TransConvDesc(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("in_channels", C.c_int), ("out_channels", C.c_int), ("kernel", C.c_int),
                ("in_padding", C.c_int), ("out_padding", C.c_int), ("flags", C.c_uint), ("leaky", C.c_float), ("clip_lo", C.c_float),
                ("clip_hi", C.c_float), ("quirks", C.c_int)]


class DwConvDesc(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("channels", C.c_int), ("downsample", C.c_int), ("dilation", C.c_int),
                ("in_padding", C.c_int), ("out_padding", C.c_int), ("flags", C.c_uint), ("leaky", C.c_float), ("clip_lo", C.c_float),
                ("clip_hi", C.c_float), ("quirks", C.c_int), ("multiplier", C.c_int), ("res_padding", C.c_int)]

# every symbol include/fyusenet_b200.h declares (tests check that the library exports all of them)
EPILOGUE_NONE, EPILOGUE_SIGMOID = 0, 1

EXPORTS = [
    "fyn_abi_version", "fyn_last_error", "fyn_device_count", "fyn_cuda_init", "fyn_cuda_shutdown",
    "fyn_get_device_info", "fyn_launch_count", "fyn_stream_create", "fyn_stream_destroy", "fyn_stream_sync",
    "fyn_event_create", "fyn_event_destroy", "fyn_event_record", "fyn_event_sync", "fyn_event_elapsed_ms",
    "fyn_stream_wait_event", "fyn_stream_add_callback", "fyn_host_alloc", "fyn_host_free", "fyn_device_alloc", "fyn_device_free",
    "fyn_memcpy_async", "fyn_download_convert", "fyn_tensor_geometry", "fyn_tensor_create",
    "fyn_tensor_wrap", "fyn_tensor_destroy", "fyn_tensor_clear", "fyn_tensor_get_desc", "fyn_tensor_device_ptr",
    "fyn_upload_f32_async", "fyn_download_f32_async", "fyn_download_f32_elems",
    "fyn_upload_u8_async", "fyn_download_u8_bytes", "fyn_download_u8_convert", "fyn_download_u8_async", "fyn_tensor_write_chw_f32",
    "fyn_tensor_read_chw_f32", "fyn_conv2d_output_size", "fyn_conv2d_create", "fyn_conv2d_load_weights",
    "fyn_conv2d_run", "fyn_conv2d_backend", "fyn_conv2d_last_kernel", "fyn_conv2d_set_epilogue", "fyn_conv2d_set_input_norm", "fyn_conv2d_plan_query",
    "fyn_conv_chain_create", "fyn_conv_chain_layers", "fyn_conv_chain_run", "fyn_conv_chain_destroy", "fyn_pool2d_create", "fyn_pool2d_run", "fyn_batchnorm_create",
    "fyn_batchnorm_load", "fyn_batchnorm_run", "fyn_sigmoid_create", "fyn_sigmoid_run", "fyn_op_destroy",
    "fyn_scale_create", "fyn_scale_out_size", "fyn_scale_run", "fyn_arith_create", "fyn_arith_run", "fyn_concat_create",
    "fyn_concat_run", "fyn_dwconv3x3_create", "fyn_dwconv3x3_load_weights", "fyn_dwconv3x3_run", "fyn_dwconv3x3_run_residual", "fyn_transconv2d_create", "fyn_transconv2d_load_weights", "fyn_transconv2d_run", "fyn_rgb2bgr_create", "fyn_rgb2bgr_run", "fyn_relayout_create", "fyn_relayout_run",
    "fyn_graph_begin_capture", "fyn_graph_end_capture", "fyn_graph_launch", "fyn_graph_destroy",
    "fyn_comm_unique_id", "fyn_comm_init", "fyn_comm_destroy", "fyn_comm_info", "fyn_allgather_logits", "fyn_comm_register_tensor", "fyn_halo_exchange",
]

_lib = None


def lib():
    """Load libfyusenet_b200.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is None:
        import os
        path = Path(os.environ.get("FYN_B200_LIB", LIB_PATH))   # (profiling builds: csrc/Makefile PROF=1)
        if not path.exists():
            raise ImportError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                              " (there is no CPU or PyTorch fallback)")
        L = C.CDLL(str(path))
        L.fyn_last_error.restype = C.c_char_p
        L.fyn_tensor_device_ptr.restype = C.c_void_p
        L.fyn_download_f32_elems.restype = C.c_size_t
        L.fyn_download_u8_bytes.restype = C.c_size_t
        for name in EXPORTS:
            getattr(L, name)
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise FynError(f"fyn status {rc}: {lib().fyn_last_error().decode(errors='replace')}")


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _s(stream):
    """cudaStream_t argument: None, a ctypes pointer, or an integer handle (e.g. torch's stream.cuda_stream)."""
    if stream is None or isinstance(stream, C.c_void_p):
        return stream
    return C.c_void_p(int(stream))


def tensor_geometry(width, height, channels, padding=0, order=ORDER_SHALLOW, dtype=F16, batch=1, packing=0):
    d = TensorDesc(width, height, channels, padding, order, dtype, batch, packing)
    g = TensorGeom()
    check(lib().fyn_tensor_geometry(C.byref(d), C.byref(g)))
    return g


class Context:
    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(lib().fyn_cuda_init(int(device), C.byref(self._h)))
        self.device = device

    def info(self) -> DeviceInfo:
        i = DeviceInfo()
        check(lib().fyn_get_device_info(self._h, C.byref(i)))
        return i

    def launch_count(self, reset=False) -> int:
        n = C.c_uint64()
        check(lib().fyn_launch_count(self._h, C.byref(n), int(reset)))
        return n.value

    def stream_create(self):
        s = C.c_void_p()
        check(lib().fyn_stream_create(self._h, C.byref(s)))
        return s

    def stream_sync(self, stream=None):
        check(lib().fyn_stream_sync(self._h, _s(stream)))

    def event_create(self):
        e = C.c_void_p()
        check(lib().fyn_event_create(self._h, C.byref(e)))
        return e

    def event_record(self, ev, stream=None):
        check(lib().fyn_event_record(self._h, ev, _s(stream)))

    def event_sync(self, ev):
        check(lib().fyn_event_sync(self._h, ev))

    def elapsed_ms(self, a, b) -> float:
        ms = C.c_float()
        check(lib().fyn_event_elapsed_ms(self._h, a, b, C.byref(ms)))
        return ms.value

    def host_alloc(self, nfloats: int) -> np.ndarray:
        """Pinned float32 host buffer as a numpy array (kept alive by the context)."""
        p = C.c_void_p()
        check(lib().fyn_host_alloc(self._h, C.c_size_t(nfloats * 4), C.byref(p)))
        buf = (C.c_float * nfloats).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.float32)
        self._pinned = getattr(self, "_pinned", []) + [(p, buf)]
        return arr

    def device_alloc(self, nbytes: int) -> int:
        """Raw device memory (fyn_device_alloc); returns the device pointer as an integer."""
        p = C.c_void_p()
        check(lib().fyn_device_alloc(self._h, C.c_size_t(int(nbytes)), C.byref(p)))
        return p.value

    def device_free(self, ptr: int):
        check(lib().fyn_device_free(self._h, C.c_void_p(int(ptr))))

    def memcpy_d2h(self, host: np.ndarray, device_ptr: int, stream=None):
        """Asynchronous device -> host copy into a (pinned) numpy array."""
        check(lib().fyn_memcpy_async(self._h, host.ctypes.data_as(C.c_void_p), C.c_void_p(int(device_ptr)), C.c_size_t(host.nbytes), 1, _s(stream)))

    def tensor(self, width, height, channels, padding=0, order=ORDER_SHALLOW, dtype=F16, batch=1, packing=0):
        return Tensor(self, TensorDesc(width, height, channels, padding, order, dtype, batch, packing))

    def close(self):
        if self._h:
            for p, _ in getattr(self, "_pinned", []):
                lib().fyn_host_free(self._h, p)
            self._pinned = []
            lib().fyn_cuda_shutdown(self._h)
            self._h = C.c_void_p()


class Tensor:
    def __init__(self, ctx: Context, desc: TensorDesc, device_ptr=None):
        self.ctx = ctx
        self._h = C.c_void_p()
        if device_ptr is None:
            check(lib().fyn_tensor_create(ctx._h, C.byref(desc), C.byref(self._h)))
        else:
            check(lib().fyn_tensor_wrap(ctx._h, C.byref(desc), C.c_void_p(device_ptr), C.byref(self._h)))
        self.desc, self.geom = TensorDesc(), TensorGeom()
        check(lib().fyn_tensor_get_desc(self._h, C.byref(self.desc), C.byref(self.geom)))

    @property
    def device_ptr(self) -> int:
        return lib().fyn_tensor_device_ptr(self._h)

    def write_chw(self, chw):
        d = self.desc
        a = np.ascontiguousarray(chw, np.float32).reshape(d.batch, d.channels, d.height, d.width)
        check(lib().fyn_tensor_write_chw_f32(self._h, _fptr(a)))

    def read_chw(self) -> np.ndarray:
        d = self.desc
        out = np.zeros((d.batch, d.channels, d.height, d.width), np.float32)
        check(lib().fyn_tensor_read_chw_f32(self._h, _fptr(out)))
        return out[0] if d.batch == 1 else out

    def upload(self, host_hwc, stream=None):
        d = self.desc
        a = np.ascontiguousarray(host_hwc, np.float32)
        assert a.size == d.batch * d.height * d.width * d.channels, (a.shape, d.batch, d.height, d.width, d.channels)
        check(lib().fyn_upload_f32_async(self._h, _fptr(a), _s(stream)))
        return a  # caller keeps it alive until the stream is synchronised

    def download_elems(self) -> int:
        return lib().fyn_download_f32_elems(self._h)

    def download(self, out=None, stream=None, sync=True) -> np.ndarray:
        g, d = self.geom, self.desc
        n = self.download_elems()
        if out is None:
            out = np.zeros(n, np.float32)
        assert out.size == n and out.dtype == np.float32
        check(lib().fyn_download_f32_async(self._h, _fptr(out), _s(stream)))
        if sync:
            self.ctx.stream_sync(stream)
        if d.order == ORDER_DEEP:
            return out.reshape(d.batch, g.tex_height, g.tex_width, 4)
        return out.reshape(d.batch, g.planes, g.tex_height, g.tex_width, 4)

    def upload_u8(self, host_hwc, stream=None):
        """UBYTE upload: uint8 [batch][H][W][C]; the tensor receives value / 255."""
        d = self.desc
        a = np.ascontiguousarray(host_hwc, np.uint8)
        assert a.size == d.batch * d.height * d.width * d.channels, (a.shape, d.batch, d.height, d.width, d.channels)
        check(lib().fyn_upload_u8_async(self._h, a.ctypes.data_as(C.POINTER(C.c_ubyte)), _s(stream)))
        return a

    def download_u8(self, stream=None) -> np.ndarray:
        """8-bit download: (uint8)(clamp(v, 0, 1) * 255) in the texel order of download()."""
        g, d = self.geom, self.desc
        out = np.zeros(lib().fyn_download_u8_bytes(self._h), np.uint8)
        check(lib().fyn_download_u8_async(self._h, out.ctypes.data_as(C.POINTER(C.c_ubyte)), _s(stream)))
        self.ctx.stream_sync(stream)
        if d.order == ORDER_DEEP:
            return out.reshape(d.batch, g.tex_height, g.tex_width, 4)
        return out.reshape(d.batch, g.planes, g.tex_height, g.tex_width, 4)

    def clear(self, stream=None):
        check(lib().fyn_tensor_clear(self._h, _s(stream)))

    def destroy(self):
        if self._h:
            lib().fyn_tensor_destroy(self._h)
            self._h = C.c_void_p()


class _Op:
    def __init__(self, ctx):
        self.ctx = ctx
        self._h = C.c_void_p()

    def destroy(self):
        if self._h:
            lib().fyn_op_destroy(self._h)
            self._h = C.c_void_p()


def conv_plan_query(stack_rows=1, **desc):
    """Device-free: the tcgen05 family's shared-memory plan for a conv descriptor (ConvDesc fields as keywords), or None if
    the family does not cover the layer."""
    d = ConvDesc()
    defaults = dict(downsample=1, dilation=1, source_step=1.0, quirks=QUIRKS_REFERENCE)
    for k, v in {**defaults, **desc}.items():
        setattr(d, k, v)
    info = ConvPlanInfo()
    rc = lib().fyn_conv2d_plan_query(C.byref(d), int(stack_rows), C.byref(info))
    return info if rc == 0 else None


def act_flags(act: str | None):
    return {None: 0, "none": 0, "relu": FLAG_PRE_RELU, "leaky": FLAG_PRE_RELU, "clip": FLAG_PRE_CLIP}[act]


class Conv2d(_Op):
    def __init__(self, ctx, weights, *, width, height, in_channels, out_channels, kernel, downsample=1, dilation=1,
                 in_padding=0, out_padding=0, res_padding=0, flags=0, leaky=0.0, clip=(0.0, 0.0), source_step=1.0,
                 fractional=False, quirks=QUIRKS_REFERENCE, backend=BACKEND_AUTO):
        super().__init__(ctx)
        self.desc = ConvDesc(width, height, in_channels, out_channels, kernel, downsample, dilation, in_padding,
                             out_padding, res_padding, flags, leaky, clip[0], clip[1], source_step,
                             int(bool(fractional)), quirks, backend)
        w = np.ascontiguousarray(weights, np.float32)
        need = out_channels + kernel * kernel * in_channels * out_channels + (
            2 * out_channels if flags & FLAG_POST_BATCHNORM else 0)
        if w.size < need:
            raise ValueError(f"weight blob too small: {w.size} < {need}")
        check(lib().fyn_conv2d_create(ctx._h, C.byref(self.desc), _fptr(w), C.byref(self._h)))
        ow, oh = C.c_int(), C.c_int()
        check(lib().fyn_conv2d_output_size(C.byref(self.desc), C.byref(ow), C.byref(oh)))
        self.out_width, self.out_height = ow.value, oh.value

    @property
    def backend(self) -> int:
        return lib().fyn_conv2d_backend(self._h)

    @property
    def last_kernel(self) -> int:
        """kernel of the last run (fyn_conv2d_last_kernel): 1 direct, 2 shallow tcgen05, 10 / 11 / 12 deep one-tile / persistent /
        halo-tile, 13 | cluster << 8 | columns << 16 deep split-K cluster kernel"""
        return lib().fyn_conv2d_last_kernel(self._h)

    def load_weights(self, weights):
        w = np.ascontiguousarray(weights, np.float32)
        check(lib().fyn_conv2d_load_weights(self._h, _fptr(w)))

    def set_input_norm(self, scale_bias):
        """Fuse the batch-norm layer in front of this (deep 1x1) convolution: scale[Cin], bias[Cin]; None switches it off."""
        if scale_bias is None:
            check(lib().fyn_conv2d_set_input_norm(self._h, None))
        else:
            sb = np.ascontiguousarray(scale_bias, np.float32)
            check(lib().fyn_conv2d_set_input_norm(self._h, _fptr(sb)))

    def set_epilogue(self, function: int):
        """Fuse the element-wise layer that follows (EPILOGUE_SIGMOID) into the convolution's epilogue."""
        check(lib().fyn_conv2d_set_epilogue(self._h, int(function)))

    def run(self, x: Tensor, out: Tensor, residual: Tensor | None = None, stream=None):
        check(lib().fyn_conv2d_run(self._h, x._h, residual._h if residual is not None else None, out._h, _s(stream)))


class ConvChain:
    """fyn_conv_chain: n >= 2 Conv2d ops of identical geometry run by one persistent kernel (bit-identical to running them one
    by one).  residual_from[i] = index of the layer whose output layer i adds (must be i - 2; -1 = the chain input)."""

    def __init__(self, ctx, ops, residual_from=None):
        self.ctx, self.ops = ctx, list(ops)
        arr = (C.c_void_p * len(self.ops))(*[o._h for o in self.ops])
        rf = None
        if residual_from is not None:
            rf = (C.c_int * len(self.ops))(*[int(v) for v in residual_from])
        self._h = C.c_void_p()
        check(lib().fyn_conv_chain_create(ctx._h, arr, rf, len(self.ops), C.byref(self._h)))

    def run(self, x, out, stream=None) -> bool:
        """False: the tensors' formats are not covered (nothing was enqueued)."""
        rc = lib().fyn_conv_chain_run(self._h, x._h, out._h, _s(stream))
        if rc == 1:
            return False
        check(rc)
        return True

    def destroy(self):
        if self._h:
            lib().fyn_conv_chain_destroy(self._h)
            self._h = C.c_void_p()


class Pool2d(_Op):
    def __init__(self, ctx, *, width, height, channels, pool=2, downsample=2, in_padding=0, out_padding=0,
                 is_max=True, global_=False, flags=0, leaky=0.0, quirks=QUIRKS_REFERENCE):
        super().__init__(ctx)
        self.desc = PoolDesc(width, height, channels, pool, pool, downsample, in_padding, out_padding,
                             int(bool(is_max)), int(bool(global_)), flags, leaky, 0.0, 0.0, quirks)
        check(lib().fyn_pool2d_create(ctx._h, C.byref(self.desc), C.byref(self._h)))

    def run(self, x, out, stream=None):
        check(lib().fyn_pool2d_run(self._h, x._h, out._h, _s(stream)))


class BatchNorm(_Op):
    def __init__(self, ctx, scale_bias, *, width, height, channels, in_padding=0, out_padding=0, flags=0):
        super().__init__(ctx)
        self.desc = BnDesc(width, height, channels, in_padding, out_padding, flags, 0.0, 0.0, 0.0)
        sb = np.ascontiguousarray(scale_bias, np.float32)
        assert sb.size >= 2 * channels
        check(lib().fyn_batchnorm_create(ctx._h, C.byref(self.desc), _fptr(sb), C.byref(self._h)))

    def run(self, x, out, stream=None):
        check(lib().fyn_batchnorm_run(self._h, x._h, out._h, _s(stream)))


class Sigmoid(_Op):
    def __init__(self, ctx, *, width, height, channels, in_padding=0, out_padding=0, flags=0):
        super().__init__(ctx)
        self.desc = UnaryDesc(width, height, channels, in_padding, out_padding, flags, 0.0, 0.0, 0.0)
        check(lib().fyn_sigmoid_create(ctx._h, C.byref(self.desc), C.byref(self._h)))

    def run(self, x, out, stream=None):
        check(lib().fyn_sigmoid_run(self._h, x._h, out._h, _s(stream)))


class DwConv3x3(_Op):
    """Depthwise 3x3 convolution (vanilla::DepthwiseConvLayer3x3 / deep::DeepDepthwiseConvLayer3x3)."""

    def __init__(self, ctx, wb, *, width, height, channels, downsample=1, dilation=1, in_padding=0, out_padding=0, flags=0,
                 leaky=0.0, quirks=None, multiplier=1, res_padding=0):
        super().__init__(ctx)
        self.desc = DwConvDesc(width, height, channels, downsample, dilation, in_padding, out_padding, flags, leaky, 0.0, 0.0,
                               QUIRKS_REFERENCE if quirks is None else quirks, multiplier, res_padding)
        self.out_width, self.out_height = width // downsample, height // downsample
        w = np.ascontiguousarray(wb, np.float32)
        check(lib().fyn_dwconv3x3_create(ctx._h, C.byref(self.desc), _fptr(w), C.byref(self._h)))

    def load_weights(self, wb):
        w = np.ascontiguousarray(wb, np.float32)
        check(lib().fyn_dwconv3x3_load_weights(self._h, _fptr(w)))

    def run(self, x, out, stream=None, residual=None):
        if residual is None:
            check(lib().fyn_dwconv3x3_run(self._h, x._h, out._h, _s(stream)))
        else:
            check(lib().fyn_dwconv3x3_run_residual(self._h, x._h, residual._h, out._h, _s(stream)))


class TransConv2d(_Op):
    """Stride-2 transpose convolution, 2x2 / 3x3, shallow (vanilla::TransConvLayer2x2 / TransConvLayer3x3)."""

    def __init__(self, ctx, wb, *, width, height, in_channels, out_channels, kernel, in_padding=0, out_padding=0, flags=0, leaky=0.0,
                 quirks=None):
        super().__init__(ctx)
        self.desc = TransConvDesc(width, height, in_channels, out_channels, kernel, in_padding, out_padding, flags, leaky, 0.0, 0.0,
                                  QUIRKS_REFERENCE if quirks is None else quirks)
        self.out_width, self.out_height = 2 * width, 2 * height
        w = np.ascontiguousarray(wb, np.float32)
        check(lib().fyn_transconv2d_create(ctx._h, C.byref(self.desc), _fptr(w), C.byref(self._h)))

    def load_weights(self, wb):
        w = np.ascontiguousarray(wb, np.float32)
        check(lib().fyn_transconv2d_load_weights(self._h, _fptr(w)))

    def run(self, x, out, stream=None):
        check(lib().fyn_transconv2d_run(self._h, x._h, out._h, _s(stream)))


class Scale(_Op):
    """ScaleLayer / DeepScaleLayer (also PADDING2D / RELU / CLIP with all factors 1)."""

    def __init__(self, ctx, *, width, height, channels, in_padding=0, out_padding=0, up=(1, 1), down=(1, 1), linear=False,
                 flags=0, leaky=0.0, clip_lo=0.0, clip_hi=0.0):
        super().__init__(ctx)
        self.desc = ScaleDesc(width, height, channels, in_padding, out_padding, up[0], up[1], down[0], down[1], int(linear),
                              flags, leaky, clip_lo, clip_hi)
        check(lib().fyn_scale_create(ctx._h, C.byref(self.desc), C.byref(self._h)))
        w, h = C.c_int(), C.c_int()
        check(lib().fyn_scale_out_size(C.byref(self.desc), C.byref(w), C.byref(h)))
        self.out_width, self.out_height = w.value, h.value

    def run(self, x, out, stream=None):
        check(lib().fyn_scale_run(self._h, x._h, out._h, _s(stream)))


class Arith(_Op):
    """AddSubLayer (two tensors, ADD / SUB) and SingletonArithmeticLayer (tensor op scalar)."""

    def __init__(self, ctx, *, width, height, channels, op, operand=None, in_padding=0, out_padding=0, flags=0):
        super().__init__(ctx)
        self.desc = ArithDesc(width, height, channels, in_padding, out_padding, op, int(operand is not None),
                              0.0 if operand is None else float(operand), flags, 0.0, 0.0, 0.0)
        check(lib().fyn_arith_create(ctx._h, C.byref(self.desc), C.byref(self._h)))

    def run(self, a, b, out, stream=None):
        check(lib().fyn_arith_run(self._h, a._h, b._h if b is not None else None, out._h, _s(stream)))


class Concat(_Op):
    def __init__(self, ctx, *, width, height, channels, in_padding=0, out_padding=0, flags=0):
        super().__init__(ctx)
        ch = (C.c_int * CONCAT_MAX_INPUTS)(*list(channels)[:CONCAT_MAX_INPUTS])
        self.desc = ConcatDesc(width, height, len(channels), ch, in_padding, out_padding, flags, 0.0, 0.0, 0.0)
        check(lib().fyn_concat_create(ctx._h, C.byref(self.desc), C.byref(self._h)))

    def run(self, inputs, out, stream=None):
        arr = (C.c_void_p * len(inputs))(*[t._h for t in inputs])
        check(lib().fyn_concat_run(self._h, arr, len(inputs), out._h, _s(stream)))


class RGB2BGR(_Op):
    def __init__(self, ctx, *, width, height, channels, in_padding=0, out_padding=0, flags=0):
        super().__init__(ctx)
        self.desc = UnaryDesc(width, height, channels, in_padding, out_padding, flags, 0.0, 0.0, 0.0)
        check(lib().fyn_rgb2bgr_create(ctx._h, C.byref(self.desc), C.byref(self._h)))

    def run(self, x, out, stream=None):
        check(lib().fyn_rgb2bgr_run(self._h, x._h, out._h, _s(stream)))


class Relayout(_Op):
    """Shallow2DeepLayer / Deep2ShallowLayer: the direction follows from the tensors' orders."""

    def __init__(self, ctx, *, width, height, channels, in_padding=0, out_padding=0, flags=0):
        super().__init__(ctx)
        self.desc = UnaryDesc(width, height, channels, in_padding, out_padding, flags, 0.0, 0.0, 0.0)
        check(lib().fyn_relayout_create(ctx._h, C.byref(self.desc), C.byref(self._h)))

    def run(self, x, out, stream=None):
        check(lib().fyn_relayout_run(self._h, x._h, out._h, _s(stream)))


COMM_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """Rank 0: the 128-byte bootstrap id (an ncclUniqueId) every rank passes to Comm()."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    check(lib().fyn_comm_unique_id(buf))
    return buf.raw


class Comm:
    """fyn_comm: NCCL communicator + CUDA-IPC peer mappings of one process-per-GPU job (include/fyusenet_b200.h, multi-GPU)."""

    def __init__(self, ctx: Context, rank: int, world: int, unique_id: bytes | None):
        self.ctx, self.rank, self.world = ctx, rank, world
        self._h = C.c_void_p()
        idbuf = C.create_string_buffer(unique_id, COMM_ID_BYTES) if unique_id is not None else None
        check(lib().fyn_comm_init(ctx._h, int(rank), int(world), idbuf, C.byref(self._h)))

    def info(self):
        r, w, v, b = C.c_int(), C.c_int(), C.c_int(), C.c_uint64()
        check(lib().fyn_comm_info(self._h, C.byref(r), C.byref(w), C.byref(v), C.byref(b)))
        return {"rank": r.value, "world": w.value, "nccl_version": v.value, "halo_bytes_pushed": b.value}

    def allgather_logits(self, logits_tensor, images_per_rank: int, device_out: int, stream=None):
        """logits_tensor: fyn_tensor* (Tensor or integer handle); device_out: device pointer to float32 [world][images_per_rank][C]."""
        h = logits_tensor._h if isinstance(logits_tensor, Tensor) else C.c_void_p(int(logits_tensor))
        check(lib().fyn_allgather_logits(self._h, h, int(images_per_rank), C.c_void_p(int(device_out)), _s(stream)))

    def register_tensor(self, tensor) -> int:
        h = tensor._h if isinstance(tensor, Tensor) else C.c_void_p(int(tensor))
        slot = C.c_int()
        check(lib().fyn_comm_register_tensor(self._h, h, C.byref(slot)))
        return slot.value

    def halo_exchange(self, slot: int, rows: int, stream=None):
        check(lib().fyn_halo_exchange(self._h, int(slot), int(rows), _s(stream)))

    def destroy(self):
        if self._h:
            check(lib().fyn_comm_destroy(self._h))
            self._h = C.c_void_p()
