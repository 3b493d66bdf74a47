"""ctypes binding of the C++ host engine's C surface (fyusenet_b200/host/capi/hostapi.cpp).

The host engine mirrors the reference's C++ API (LayerBuilder / LayerFactory / NeuralNetwork / BufferManager /
the StyleNet and ResNet-50 sample networks); this module lets Python (tests, bench.py) drive exactly the
call sequence of samples/desktop/stylenet.cpp:148-187 / resnet.cpp:138-160:
loadWeightsAndBiases -> setup -> setInputBuffer -> forward -> getOutputBuffer.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import capi

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libfyusenet_host.so"
_lib = None


class HostError(RuntimeError):
    """A FynException caught at the C surface."""


def lib():
    global _lib
    if _lib is None:
        capi.lib()  # libfyusenet_b200.so first (rpath $ORIGIN resolves it as well)
        if not LIB_PATH.exists():
            raise ImportError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
        L = C.CDLL(str(LIB_PATH))
        L.fynhost_last_error.restype = C.c_char_p
        L.fynhost_stylenet_create.restype = C.c_void_p
        L.fynhost_resnet50_create.restype = C.c_void_p
        L.fynhost_layerzoo_create.restype = C.c_void_p
        L.fynhost_net_weight_floats.restype = C.c_size_t
        L.fynhost_net_weight_offset.restype = C.c_longlong
        L.fynhost_net_output.restype = C.POINTER(C.c_float)
        L.fynhost_net_input_buffer.restype = C.POINTER(C.c_float)
        L.fynhost_stylenet_input_buffer_slot.restype = C.POINTER(C.c_float)
        L.fynhost_net_async_completed.restype = C.c_uint64
        L.fynhost_stylenet_output_tensor.restype = C.c_void_p
        L.fynhost_net_context.restype = C.c_void_p
        L.fynhost_net_stream.restype = C.c_void_p
        L.fynhost_net_device_bytes.restype = C.c_size_t
        L.fynhost_net_layer_tensor.restype = C.c_void_p
        L.fynhost_net_input_raw.restype = C.c_void_p
        L.fynhost_net_output_raw.restype = C.c_void_p
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise HostError(lib().fynhost_last_error().decode(errors="replace"))


def async_slots() -> int:
    """Buffers per pipeline interface == max sequences in flight of the asynchronous engine."""
    return lib().fynhost_async_slots()


def selftest():
    buf = C.create_string_buffer(16384)
    n = lib().fynhost_selftest(buf, len(buf))
    return n, buf.value.decode()


def set_storage_precision(fp32: bool):
    _check(lib().fynhost_set_storage_precision(int(bool(fp32))))


class Network:
    """Common driver for the sample networks."""

    def __init__(self, handle):
        if not handle:
            raise HostError(lib().fynhost_last_error().decode(errors="replace"))
        self._h = C.c_void_p(handle)
        self._setup = False

    # -- weights -------------------------------------------------------------------------------
    @property
    def weight_floats(self) -> int:
        return lib().fynhost_net_weight_floats(self._h)

    def weight_offset(self, layer_number: int) -> int:
        return lib().fynhost_net_weight_offset(self._h, int(layer_number))

    def load_weights(self, w):
        w = np.ascontiguousarray(w, np.float32)
        _check(lib().fynhost_net_load_weights(self._h, w.ctypes.data_as(C.POINTER(C.c_float)), C.c_size_t(w.size)))

    # -- lifecycle -----------------------------------------------------------------------------
    def set_batch(self, n: int):
        _check(lib().fynhost_net_set_batch(self._h, int(n)))

    def setup(self):
        _check(lib().fynhost_net_setup(self._h))
        self._setup = True

    def set_input(self, hwc):
        a = np.ascontiguousarray(hwc, np.float32)
        _check(lib().fynhost_net_set_input(self._h, a.ctypes.data_as(C.POINTER(C.c_float))))

    def input_buffer(self) -> np.ndarray:
        """The network's pinned upload buffer as a numpy view: fill it in place, then forward()."""
        n = C.c_size_t()
        p = lib().fynhost_net_input_buffer(self._h, C.byref(n))
        if not p:
            raise HostError(lib().fynhost_last_error().decode(errors="replace"))
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def forward(self):
        _check(lib().fynhost_net_forward(self._h))

    def finish(self):
        _check(lib().fynhost_net_finish(self._h))

    # -- asynchronous (pipelined) operation ---------------------------------------------------
    def asynchronous(self):
        """NeuralNetwork::asynchronous(): call before setup(); forward() then only enqueues (<= async_slots() sequences in flight)."""
        _check(lib().fynhost_net_asynchronous(self._h))

    def async_completed(self):
        """(number of delivered sequences, last delivered sequence number, pointer to its host buffer or None)"""
        seq, data = C.c_uint64(), C.POINTER(C.c_float)()
        n = lib().fynhost_net_async_completed(self._h, C.byref(seq), C.byref(data))
        return int(n), int(seq.value), data

    def output(self) -> np.ndarray:
        n = C.c_size_t()
        p = lib().fynhost_net_output(self._h, C.byref(n))
        if not p:
            raise HostError(lib().fynhost_last_error().decode(errors="replace"))
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def use_stream(self, stream):
        _check(lib().fynhost_net_use_stream(self._h, capi._s(stream)))

    @property
    def stream(self):
        return C.c_void_p(lib().fynhost_net_stream(self._h))

    # -- introspection -------------------------------------------------------------------------
    def layers(self):
        out = []
        for i in range(lib().fynhost_net_num_layers(self._h)):
            no, ch, w, h, fam = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
            name = C.create_string_buffer(128)
            _check(lib().fynhost_net_layer_info(self._h, i, C.byref(no), C.byref(ch), C.byref(w), C.byref(h), C.byref(fam), name, 128))
            out.append(dict(number=no.value, name=name.value.decode(), channels=ch.value, width=w.value, height=h.value,
                            family=fam.value))
        return out

    def layer_result(self, number: int, shape) -> np.ndarray:
        out = np.zeros(shape, np.float32)
        _check(lib().fynhost_net_copy_layer_result(self._h, int(number), out.ctypes.data_as(C.POINTER(C.c_float)),
                                                   C.c_size_t(out.size)))
        return out

    def enable_dumps(self, directory: str):
        _check(lib().fynhost_net_enable_dumps(self._h, str(directory).encode()))

    def enable_fusion(self, on=True):
        """Engine-level layer fusion (conv + sigmoid in one kernel); on by default, suspended while dumps are written."""
        _check(lib().fynhost_net_enable_fusion(self._h, int(on)))

    @property
    def fused_layers(self) -> int:
        return lib().fynhost_net_fused_layers(self._h)

    def enable_chains(self, on=True):
        """Runs of same-geometry convolutions as one persistent kernel (Engine::enableChains); on by default."""
        _check(lib().fynhost_net_enable_chains(self._h, int(bool(on))))

    @property
    def chained_layers(self) -> int:
        return lib().fynhost_net_chained_layers(self._h)

    @property
    def halo_exchanges(self) -> int:
        """Exchanges per forward in row-banded operation (Engine::planHalo): margins are refreshed only where a layer would reach spoilt rows."""
        return lib().fynhost_net_halo_exchanges(self._h)

    def enable_timings(self, on=True):
        _check(lib().fynhost_net_enable_timings(self._h, int(on)))

    def enable_layer_timing(self, number: int):
        """Event pair around one layer only (the others keep their dependent-launch overlap); read with layer_timing()."""
        _check(lib().fynhost_net_enable_layer_timing(self._h, int(number)))

    def layer_timing(self, number: int):
        ms, us = C.c_float(), C.c_uint()
        _check(lib().fynhost_net_layer_timing(self._h, int(number), C.byref(ms), C.byref(us)))
        return ms.value, us.value

    def enable_graph(self, on=True):
        """Replay the device layers of the synchronous path from a CUDA graph (Engine::enableGraph)."""
        _check(lib().fynhost_net_enable_graph(self._h, int(bool(on))))

    @property
    def graph_active(self) -> bool:
        return bool(lib().fynhost_net_graph_active(self._h))

    def skip_io(self, on=True):
        """Skip the upload / download layers: the input of the last real upload stays resident in HBM, the result stays in
        the last layer's tensor (device-resident timing of a network built with I/O layers)."""
        _check(lib().fynhost_net_skip_io(self._h, int(bool(on))))

    def layer_tensor(self, number: int) -> int:
        """fyn_tensor* (as an integer) of a layer's output tensor."""
        t = lib().fynhost_net_layer_tensor(self._h, int(number))
        if not t:
            raise HostError(lib().fynhost_last_error().decode(errors="replace"))
        return t

    def set_halo_exchange(self, comm, margin_rows: int, input_height: int):
        """Row-banded operation: `comm` is a capi.Comm (or None to switch it off); collective."""
        _check(lib().fynhost_net_set_halo_exchange(self._h, comm._h if comm is not None else None, int(margin_rows), int(input_height)))

    def device_bytes(self) -> int:
        return lib().fynhost_net_device_bytes(self._h)

    @property
    def num_tensors(self) -> int:
        return lib().fynhost_net_num_tensors(self._h)

    def destroy(self):
        if self._h:
            lib().fynhost_net_destroy(self._h)
            self._h = C.c_void_p()


class StyleNet(Network):
    def __init__(self, kernel: int, width: int, height: int, upload=True, download=True, device=0):
        super().__init__(lib().fynhost_stylenet_create(int(kernel), int(width), int(height), int(upload), int(download), int(device)))
        self.kernel, self.width, self.height = kernel, width, height
        self.byte_io = False

    def set_byte_io(self, on=True):
        """8-bit frames in and out (StyleNetBase::setByteIO, before setup()): UBYTE upload (value / 255 on the device) and RGBA8
        download ((uint8)(clamp(v, 0, 1) * 255) on the device); input_buffer() / input_buffer_slot() / output_rgba() are uint8."""
        _check(lib().fynhost_stylenet_set_byte_io(self._h, int(bool(on))))
        self.byte_io = bool(on)

    def _raw(self, fn, *args):
        n, dt = C.c_size_t(), C.c_int()
        p = fn(self._h, *args, C.byref(n), C.byref(dt))
        if not p:
            raise HostError(lib().fynhost_last_error().decode(errors="replace"))
        ctype, count = {0: (C.c_float, n.value // 4), 1: (C.c_uint16, n.value // 2), 2: (C.c_ubyte, n.value)}[dt.value]
        return np.ctypeslib.as_array(C.cast(C.c_void_p(p), C.POINTER(ctype)), shape=(count,))

    def input_buffer(self) -> np.ndarray:
        return self._raw(lib().fynhost_net_input_raw, -1)

    def output(self) -> np.ndarray:
        return self._raw(lib().fynhost_net_output_raw)

    def set_input_tensor(self, tensor: capi.Tensor):
        _check(lib().fynhost_stylenet_set_input_tensor(self._h, tensor._h))

    def input_buffer_slot(self, slot: int) -> np.ndarray:
        """Pinned input buffer `slot` of an asynchronous network (sequence s reads slot s % async_slots())."""
        return self._raw(lib().fynhost_net_input_raw, int(slot))

    def output_rgba(self) -> np.ndarray:
        """download buffer as [batch?][H][W][4] float32 -- uint8 with set_byte_io() -- (RGBA, alpha = 0.5: compare RGB only)."""
        return self.output().reshape(-1, self.height, self.width, 4)


class ResNet50(Network):
    def __init__(self, device=0, batch=1):
        super().__init__(lib().fynhost_resnet50_create(int(device), int(batch)))
        self.batch = batch
        self.byte_input = False

    def set_byte_input(self, on=True):
        """8-bit images in (ResNet50::setByteInput, before setup()): UBYTE upload, value / 255 on the device -- the conversion the
        reference's sample does on the host (samples/desktop/resnet.cpp:48-51), bit for bit; 3 instead of 12 bytes per pixel over PCIe."""
        _check(lib().fynhost_resnet50_set_byte_input(self._h, int(bool(on))))
        self.byte_input = bool(on)

    def input_buffer(self) -> np.ndarray:
        if not self.byte_input:
            return super().input_buffer()
        n, dt = C.c_size_t(), C.c_int()
        lib().fynhost_net_input_raw.restype = C.c_void_p
        p = lib().fynhost_net_input_raw(self._h, -1, C.byref(n), C.byref(dt))
        if not p:
            raise HostError(lib().fynhost_last_error().decode(errors="replace"))
        return np.ctypeslib.as_array(C.cast(C.c_void_p(p), C.POINTER(C.c_ubyte)), shape=(n.value,))

    def logits(self) -> np.ndarray:
        """[batch][1000]: the deep 18x14 download texture is channel order for 1x1 spatial (cpubuffer.cpp:131-142)."""
        return self.output().reshape(self.batch, -1)[:, :1000]


class LayerZoo(Network):
    """samplenetworks/layerzoo.h: upload -> rgb2bgr -> scale x2 (linear) -> *2 -> sub -> concat -> clip -> shallow2deep ->
    deep scale /2 -> deep2shallow -> padding -> add -> download; no weights."""
    LAYERS = ("upload", "bgr", "upscale", "twice", "diff", "concat", "clip", "todeep", "downscale", "toshallow", "pad", "sum",
              "download")
    CLIP = (0.2, 0.7)

    def __init__(self, width: int, height: int, device=0):
        super().__init__(lib().fynhost_layerzoo_create(int(width), int(height), int(device)))
        self.width, self.height = width, height
