"""ctypes binding of the OSPRay C API (include/ospray/ospray.h, ospray_util.h) plus the cuda-device
extensions (include/ospray_cuda_ext.h). Mirrors the reference's header-only C++ wrappers
(ospray/include/ospray/ospray_cpp/*.h) closely enough that scene scripts read like ospray_testing."""
import ctypes as C
import numpy as np
from . import lib_path

# ---- enums (include/ospray/OSPEnums.h) -----------------------------------------------------------
OSP_DEVICE, OSP_VOID_PTR, OSP_BOOL = 100, 200, 250
OSP_OBJECT = 0x8000000
(OSP_DATA, OSP_CAMERA, OSP_FRAMEBUFFER, OSP_FUTURE, OSP_GEOMETRIC_MODEL, OSP_GEOMETRY, OSP_GROUP, OSP_IMAGE_OPERATION,
 OSP_INSTANCE, OSP_LIGHT, OSP_MATERIAL, OSP_RENDERER, OSP_TEXTURE, OSP_TRANSFER_FUNCTION, OSP_VOLUME,
 OSP_VOLUMETRIC_MODEL, OSP_WORLD) = range(OSP_OBJECT + 100, OSP_OBJECT + 117)
OSP_STRING = 1500
OSP_UCHAR, OSP_VEC2UC, OSP_VEC3UC, OSP_VEC4UC = 2500, 2501, 2502, 2503
OSP_USHORT, OSP_VEC2US, OSP_VEC3US, OSP_VEC4US = 3500, 3501, 3502, 3503
OSP_HALF, OSP_VEC2H, OSP_VEC3H, OSP_VEC4H = 5800, 5801, 5802, 5803
OSP_LINEAR2F = 12000
# OSPTextureFormat / OSPTextureFilter / OSPTextureWrapMode
OSP_TEXTURE_FORMATS = ["RGBA8", "SRGBA", "RGBA32F", "RGB8", "SRGB", "RGB32F", "R8", "R32F", "L8", "RA8", "LA8", "RGBA16", "RGB16", "RA16", "R16",
                       "RA32F", "RGBA16F", "RGB16F", "RA16F", "R16F"]
OSP_TEXTURE_FILTER_LINEAR, OSP_TEXTURE_FILTER_NEAREST = 0, 1
OSP_TEXTURE_WRAP_REPEAT, OSP_TEXTURE_WRAP_MIRRORED_REPEAT, OSP_TEXTURE_WRAP_CLAMP_TO_EDGE = 0, 1, 2
OSP_INT, OSP_VEC2I, OSP_VEC3I, OSP_VEC4I = 4000, 4001, 4002, 4003
OSP_UINT, OSP_VEC2UI, OSP_VEC3UI, OSP_VEC4UI = 4500, 4501, 4502, 4503
OSP_FLOAT, OSP_VEC2F, OSP_VEC3F, OSP_VEC4F = 6000, 6001, 6002, 6003
OSP_BOX1F, OSP_BOX2F, OSP_BOX3F = 10000, 10001, 10002
OSP_LINEAR3F, OSP_AFFINE3F = 12001, 12003
OSP_FB_NONE, OSP_FB_RGBA8, OSP_FB_SRGBA, OSP_FB_RGBA32F = 0, 1, 2, 3
OSP_FB_COLOR, OSP_FB_DEPTH, OSP_FB_ACCUM, OSP_FB_VARIANCE = 1, 2, 4, 8
OSP_FB_NORMAL, OSP_FB_ALBEDO, OSP_FB_ID_PRIMITIVE, OSP_FB_ID_OBJECT, OSP_FB_ID_INSTANCE = 16, 32, 64, 128, 256
OSP_FB_FIRST_NORMAL, OSP_FB_POSITION = 512, 1024
OSP_PIXELFILTER_POINT, OSP_PIXELFILTER_BOX, OSP_PIXELFILTER_GAUSS = 0, 1, 2
OSP_NONE_FINISHED, OSP_WORLD_RENDERED, OSP_FRAME_FINISHED, OSP_TASK_FINISHED = 0, 10, 30, 100000
OSP_NO_ERROR = 0

_lib = None
_ext = None
_errors = []


class OSPBounds(C.Structure):
    _fields_ = [("lower", C.c_float * 3), ("upper", C.c_float * 3)]


_ERRCB = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_char_p)
_MSGCB = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p)


@_ERRCB
def _on_error(_ud, code, msg):
    _errors.append((int(code), (msg or b"").decode(errors="replace")))


def lib():
    """libospray.so, loaded once with RTLD_GLOBAL so that the module resolves against it."""
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(lib_path("libospray.so"), mode=C.RTLD_GLOBAL)
    P, S, U32, U64, I64, F = C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint64, C.c_int64, C.c_float
    sig = {
        "ospInit": (U32, [C.POINTER(C.c_int), C.POINTER(S)]), "ospShutdown": (None, []),
        "ospLoadModule": (U32, [S]), "ospNewDevice": (P, [S]), "ospSetCurrentDevice": (None, [P]),
        "ospGetCurrentDevice": (P, []), "ospDeviceSetParam": (None, [P, S, U32, P]), "ospDeviceCommit": (None, [P]),
        "ospDeviceRelease": (None, [P]), "ospDeviceRetain": (None, [P]), "ospDeviceRemoveParam": (None, [P, S]),
        "ospDeviceGetProperty": (I64, [P, U32]), "ospDeviceSetErrorCallback": (None, [P, _ERRCB, P]),
        "ospDeviceSetStatusCallback": (None, [P, _MSGCB, P]), "ospDeviceGetLastErrorCode": (U32, [P]),
        "ospDeviceGetLastErrorMsg": (S, [P]),
        "ospNewSharedData": (P, [P, U32, U64, I64, U64, I64, U64, I64, P, P]), "ospNewData": (P, [U32, U64, U64, U64]),
        "ospCopyData": (None, [P, P, U64, U64, U64]),
        "ospNewLight": (P, [S]), "ospNewCamera": (P, [S]), "ospNewGeometry": (P, [S]), "ospNewVolume": (P, [S]),
        "ospNewGeometricModel": (P, [P]), "ospNewVolumetricModel": (P, [P]), "ospNewMaterial": (P, [S]),
        "ospNewTransferFunction": (P, [S]), "ospNewTexture": (P, [S]), "ospNewGroup": (P, []), "ospNewInstance": (P, [P]),
        "ospNewWorld": (P, []), "ospGetBounds": (OSPBounds, [P]),
        "ospSetParam": (None, [P, S, U32, P]), "ospRemoveParam": (None, [P, S]), "ospCommit": (None, [P]),
        "ospRelease": (None, [P]), "ospRetain": (None, [P]),
        "ospNewFrameBuffer": (P, [C.c_int, C.c_int, U32, U32]), "ospNewImageOperation": (P, [S]),
        "ospMapFrameBuffer": (P, [P, U32]), "ospUnmapFrameBuffer": (None, [P, P]), "ospGetVariance": (F, [P]),
        "ospResetAccumulation": (None, [P]), "ospNewRenderer": (P, [S]), "ospRenderFrame": (P, [P, P, P, P]),
        "ospIsReady": (C.c_int, [P, U32]), "ospWait": (None, [P, U32]), "ospCancel": (None, [P]),
        "ospGetProgress": (F, [P]), "ospGetTaskDuration": (F, [P]), "ospRenderFrameBlocking": (F, [P, P, P, P]),
        "ospPick": (None, [P, P, P, P, P, F, F]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def ext():
    """cuda-device extensions exported by the module .so (available after ospLoadModule('cuda'))."""
    global _ext
    if _ext is None:
        E = C.CDLL(lib_path("libospray_module_cuda.so"), mode=C.RTLD_GLOBAL)
        E.ospcudaFrameBufferDevicePointer.restype = C.c_void_p
        E.ospcudaFrameBufferDevicePointer.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)]
        E.ospcudaGetFusedRayCount.restype = C.c_uint64
        E.ospcudaGetFusedRayCount.argtypes = [C.c_void_p]
        E.ospcudaGetRayCounts.restype = None
        E.ospcudaGetRayCounts.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        E.ospcudaGetGroupBuildTime.restype = C.c_float
        E.ospcudaGetGroupBuildTime.argtypes = [C.c_void_p]
        E.ospcudaKernelLaunchCount.restype = C.c_uint64
        E.ospcudaRenderStream.restype = C.c_void_p
        E.ospcudaWorldDevicePointer.restype = C.c_void_p
        E.ospcudaGetFrameStageTimes.restype = None
        E.ospcudaGetFrameStageTimes.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        E.ospcudaSetRenderStream.restype = None
        E.ospcudaSetRenderStream.argtypes = [C.c_void_p]
        E.ospcudaWorldDevicePointer.argtypes = [C.c_void_p]
        E.ospcudaTextureDevicePointer.restype = C.c_void_p
        E.ospcudaTextureDevicePointer.argtypes = [C.c_void_p]
        _ext = E
    return _ext


class OSPRayError(RuntimeError):
    pass


def check_errors():
    if _errors:
        errs = list(_errors)
        _errors.clear()
        raise OSPRayError("; ".join(f"[{c}] {m}" for c, m in errs))


def pending_errors():
    errs = list(_errors)
    _errors.clear()
    return errs


_device = None


def init(device="cuda", module="cuda", warn_as_error=True, **device_params):
    """ospLoadModule + ospNewDevice + params + commit + ospSetCurrentDevice (API.cpp:171-187 path)."""
    global _device
    L = lib()
    if _device is not None:
        return _device
    if L.ospLoadModule(module.encode()) != OSP_NO_ERROR:
        raise OSPRayError(f"ospLoadModule({module}) failed")
    d = L.ospNewDevice(device.encode())
    if not d:
        raise OSPRayError(f"ospNewDevice({device}) failed")
    L.ospDeviceSetErrorCallback(d, _on_error, None)
    if warn_as_error:
        b = C.c_bool(True)
        L.ospDeviceSetParam(d, b"warnAsError", OSP_BOOL, C.byref(b))
    for k, v in device_params.items():
        iv = C.c_int(int(v))
        L.ospDeviceSetParam(d, k.encode(), OSP_INT, C.byref(iv))
    L.ospDeviceCommit(d)
    code = L.ospDeviceGetLastErrorCode(d)
    if code != OSP_NO_ERROR:
        msg = L.ospDeviceGetLastErrorMsg(d)
        L.ospDeviceRelease(d)
        _errors.clear()
        raise OSPRayError(f"device commit failed: {(msg or b'').decode()}")
    L.ospSetCurrentDevice(d)
    _device = d
    check_errors()
    return d


def new_device(device="cuda", warn_as_error=True, **device_params):
    """an additional device object (e.g. one per tile-partition rank); make it current with set_current()"""
    L = lib()
    d = L.ospNewDevice(device.encode())
    if not d:
        raise OSPRayError(f"ospNewDevice({device}) failed")
    L.ospDeviceSetErrorCallback(d, _on_error, None)
    if warn_as_error:
        b = C.c_bool(True)
        L.ospDeviceSetParam(d, b"warnAsError", OSP_BOOL, C.byref(b))
    for k, v in device_params.items():
        iv = C.c_int(int(v))
        L.ospDeviceSetParam(d, k.encode(), OSP_INT, C.byref(iv))
    L.ospDeviceCommit(d)
    if L.ospDeviceGetLastErrorCode(d) != OSP_NO_ERROR:
        msg = L.ospDeviceGetLastErrorMsg(d)
        L.ospDeviceRelease(d)
        _errors.clear()
        raise OSPRayError(f"device commit failed: {(msg or b'').decode()}")
    return d


def set_device_bool(name, value):
    """ospDeviceSetParam + ospDeviceCommit of a bool parameter on the current device (e.g. disableMipMapGeneration)"""
    b = C.c_bool(bool(value))
    lib().ospDeviceSetParam(_device, name.encode(), OSP_BOOL, C.byref(b))
    lib().ospDeviceCommit(_device)
    check_errors()


def set_current(device):
    global _device
    lib().ospSetCurrentDevice(device)
    _device = device


def shutdown():
    global _device
    if _device is not None:
        lib().ospShutdown()
        lib().ospDeviceRelease(_device)
        _device = None


# ---- typed scalars ------------------------------------------------------------------------------
class Typed:
    """explicitly typed parameter value: Typed(OSP_UINT, 3), Typed(OSP_BOX1F, (0, 10))"""

    def __init__(self, osp_type, value):
        self.type, self.value = osp_type, value


def _pack(value):
    """python value -> (OSPDataType, ctypes buffer)"""
    if isinstance(value, Typed):
        t, v = value.type, value.value
        if t in (OSP_UINT,):
            return t, C.c_uint32(int(v))
        if t == OSP_INT:
            return t, C.c_int(int(v))
        if t == OSP_BOOL:
            return t, C.c_bool(bool(v))
        arr = np.ascontiguousarray(np.asarray(v, dtype=np.int32 if t in (OSP_VEC2I, OSP_VEC3I, OSP_VEC4I)
                                              else np.uint32 if t in (OSP_VEC2UI, OSP_VEC3UI, OSP_VEC4UI) else np.float32))
        return t, (C.c_char * arr.nbytes).from_buffer_copy(arr.tobytes())
    if isinstance(value, bool):
        return OSP_BOOL, C.c_bool(value)
    if isinstance(value, int):
        return OSP_INT, C.c_int(value)
    if isinstance(value, float):
        return OSP_FLOAT, C.c_float(value)
    if isinstance(value, (tuple, list)) and all(isinstance(x, (int, float)) for x in value):
        n = len(value)
        if all(isinstance(x, int) for x in value):
            t = {2: OSP_VEC2I, 3: OSP_VEC3I, 4: OSP_VEC4I}[n]
            return t, (C.c_int * n)(*value)
        t = {2: OSP_VEC2F, 3: OSP_VEC3F, 4: OSP_VEC4F}[n]
        return t, (C.c_float * n)(*[float(x) for x in value])
    raise TypeError(f"cannot infer OSPDataType for {value!r}")


_NP_TYPES = {
    (np.dtype(np.float32), 1): OSP_FLOAT, (np.dtype(np.float32), 2): OSP_VEC2F, (np.dtype(np.float32), 3): OSP_VEC3F,
    (np.dtype(np.float32), 4): OSP_VEC4F, (np.dtype(np.uint32), 1): OSP_UINT, (np.dtype(np.uint32), 3): OSP_VEC3UI,
    (np.dtype(np.uint32), 4): OSP_VEC4UI, (np.dtype(np.uint8), 1): OSP_UCHAR, (np.dtype(np.float32), 6): OSP_BOX3F,
}


class Object:
    """handle wrapper: releases on garbage collection (like cpp::ManagedObject)"""
    kind = OSP_OBJECT

    def __init__(self, handle):
        if not handle:
            errs = pending_errors()
            raise OSPRayError(f"{type(self).__name__}: creation failed {errs}")
        self.h = handle
        self._keep = []

    def __del__(self):
        try:
            if self.h and _lib is not None and _device is not None:
                _lib.ospRelease(self.h)
        except Exception:
            pass

    def set(self, name, value):
        L = lib()
        if isinstance(value, Object):
            hv = C.c_void_p(value.h)
            L.ospSetParam(self.h, name.encode(), value.kind, C.byref(hv))
        elif isinstance(value, np.ndarray):
            d = Data(value)
            hv = C.c_void_p(d.h)
            L.ospSetParam(self.h, name.encode(), OSP_DATA, C.byref(hv))
        elif isinstance(value, (list, tuple)) and value and isinstance(value[0], Object):
            d = Data.of_objects(value)
            hv = C.c_void_p(d.h)
            L.ospSetParam(self.h, name.encode(), OSP_DATA, C.byref(hv))
        elif isinstance(value, str):
            L.ospSetParam(self.h, name.encode(), OSP_STRING, value.encode())
        else:
            t, buf = _pack(value)
            L.ospSetParam(self.h, name.encode(), t, C.byref(buf))
        return self

    def remove(self, name):
        lib().ospRemoveParam(self.h, name.encode())

    def commit(self):
        lib().ospCommit(self.h)
        check_errors()
        return self

    def bounds(self):
        b = lib().ospGetBounds(self.h)
        return np.array(list(b.lower), dtype=np.float32), np.array(list(b.upper), dtype=np.float32)


class Data(Object):
    kind = OSP_DATA

    def __init__(self, array, osp_type=None, dims=None, shared=False):
        """copied data (cpp::CopiedData, ospray_cpp/Data.h:227-241): ospNewSharedData -> ospNewData ->
        ospCopyData -> release(tmp); shared=True keeps the numpy buffer alive instead."""
        L = lib()
        a = np.ascontiguousarray(array)
        if osp_type is None:
            if dims is None and a.ndim == 3 and a.dtype == np.float32:  # 3-D scalar field (z, y, x)
                osp_type, dims = OSP_FLOAT, (a.shape[2], a.shape[1], a.shape[0])
            else:
                comps = 1 if a.ndim == 1 else a.shape[-1]
                osp_type = _NP_TYPES[(a.dtype, comps)]
        if dims is None:
            dims = (a.shape[0], 1, 1)
        self.array = a
        tmp = L.ospNewSharedData(a.ctypes.data_as(C.c_void_p), osp_type, dims[0], 0, dims[1], 0, dims[2], 0, None, None)
        if shared:
            super().__init__(tmp)
            return
        h = L.ospNewData(osp_type, dims[0], dims[1], dims[2])
        L.ospCopyData(tmp, h, 0, 0, 0)
        L.ospRelease(tmp)
        super().__init__(h)
        L.ospCommit(h)

    @staticmethod
    def of_objects(objs):
        L = lib()
        kind = objs[0].kind
        arr = (C.c_void_p * len(objs))(*[o.h for o in objs])
        d = Data.__new__(Data)
        tmp = L.ospNewSharedData(C.cast(arr, C.c_void_p), kind, len(objs), 0, 1, 0, 1, 0, None, None)
        h = L.ospNewData(kind, len(objs), 1, 1)
        L.ospCopyData(tmp, h, 0, 0, 0)
        L.ospRelease(tmp)
        Object.__init__(d, h)
        d.array = arr
        L.ospCommit(h)
        return d


def _typed_object(name, kind, ctor):
    def __init__(self, type_name):
        Object.__init__(self, getattr(lib(), ctor)(type_name.encode()))
    return type(name, (Object,), {"kind": kind, "__init__": __init__})


Light = _typed_object("Light", OSP_LIGHT, "ospNewLight")
Camera = _typed_object("Camera", OSP_CAMERA, "ospNewCamera")
Geometry = _typed_object("Geometry", OSP_GEOMETRY, "ospNewGeometry")
Volume = _typed_object("Volume", OSP_VOLUME, "ospNewVolume")
Material = _typed_object("Material", OSP_MATERIAL, "ospNewMaterial")
TransferFunction = _typed_object("TransferFunction", OSP_TRANSFER_FUNCTION, "ospNewTransferFunction")
Renderer = _typed_object("Renderer", OSP_RENDERER, "ospNewRenderer")
ImageOperation = _typed_object("ImageOperation", OSP_IMAGE_OPERATION, "ospNewImageOperation")
Texture = _typed_object("Texture", OSP_TEXTURE, "ospNewTexture")


def texture2d(array, fmt, filter=OSP_TEXTURE_FILTER_LINEAR, wrap=(OSP_TEXTURE_WRAP_REPEAT, OSP_TEXTURE_WRAP_REPEAT)):
    """texture2d from a (h, w[, channels]) array of uint8 / uint16 / float16 / float32; `fmt` names an OSPTextureFormat"""
    a = np.ascontiguousarray(array)
    h, w = a.shape[:2]
    comps = 1 if a.ndim == 2 else a.shape[2]
    base = {np.dtype(np.uint8): OSP_UCHAR, np.dtype(np.uint16): OSP_USHORT, np.dtype(np.float16): OSP_HALF, np.dtype(np.float32): OSP_FLOAT}[a.dtype]
    t = Texture("texture2d")
    t.set("format", Typed(OSP_UINT, OSP_TEXTURE_FORMATS.index(fmt)))
    t.set("filter", Typed(OSP_UINT, filter))
    t.set("wrapMode", Typed(OSP_VEC2UI, wrap))
    d = Data(a, osp_type=base + comps - 1, dims=(w, h, 1))
    hv = C.c_void_p(d.h)
    lib().ospSetParam(t.h, b"data", OSP_DATA, C.byref(hv))
    t._keep.append(d)
    return t.commit()


class GeometricModel(Object):
    kind = OSP_GEOMETRIC_MODEL

    def __init__(self, geometry):
        super().__init__(lib().ospNewGeometricModel(geometry.h))
        self._keep.append(geometry)


class VolumetricModel(Object):
    kind = OSP_VOLUMETRIC_MODEL

    def __init__(self, volume):
        super().__init__(lib().ospNewVolumetricModel(volume.h))
        self._keep.append(volume)


class Group(Object):
    kind = OSP_GROUP

    def __init__(self):
        super().__init__(lib().ospNewGroup())

    def build_ms(self):
        return float(ext().ospcudaGetGroupBuildTime(self.h))


class Instance(Object):
    kind = OSP_INSTANCE

    def __init__(self, group):
        super().__init__(lib().ospNewInstance(group.h))
        self._keep.append(group)


class World(Object):
    kind = OSP_WORLD

    def __init__(self):
        super().__init__(lib().ospNewWorld())

    def device_pointer(self):
        """const OcWorld* in device memory (for ospcuda_trace)"""
        return ext().ospcudaWorldDevicePointer(self.h)


class PickResult(C.Structure):
    _fields_ = [("hasHit", C.c_int), ("worldPosition", C.c_float * 3), ("instance", C.c_void_p), ("model", C.c_void_p), ("primID", C.c_uint32)]


def pick(fb, renderer, camera, world, x, y):
    """ospPick at normalised screen position (x, y); returns (hasHit, worldPosition, instance handle, model handle, primID).
    The returned handles are retained by the device (Renderer.cpp:128-129) and released here after reading their values."""
    L = lib()
    r = PickResult()
    L.ospPick(C.byref(r), fb.h, renderer.h, camera.h, world.h, float(x), float(y))
    check_errors()
    out = (bool(r.hasHit), np.array(list(r.worldPosition), np.float32), r.instance, r.model, int(r.primID))
    if r.hasHit:
        L.ospRelease(r.instance)
        L.ospRelease(r.model)
    return out


class Future(Object):
    kind = OSP_FUTURE

    def wait(self, event=OSP_TASK_FINISHED):
        lib().ospWait(self.h, event)
        check_errors()

    def is_ready(self, event=OSP_TASK_FINISHED):
        return bool(lib().ospIsReady(self.h, event))

    def duration(self):
        return float(lib().ospGetTaskDuration(self.h))

    def cancel(self):
        lib().ospCancel(self.h)
        check_errors()

    def progress(self):
        L = lib()
        L.ospGetProgress.restype = C.c_float
        return float(L.ospGetProgress(self.h))

    def stage_ms(self):
        """(renderer kernel ms, post-process ms) of the finished frame"""
        ms = (C.c_float * 2)()
        ext().ospcudaGetFrameStageTimes(self.h, ms)
        return float(ms[0]), float(ms[1])


_CHANNEL_DTYPE = {
    OSP_FB_DEPTH: (np.float32, 1), OSP_FB_NORMAL: (np.float32, 3), OSP_FB_ALBEDO: (np.float32, 3),
    OSP_FB_FIRST_NORMAL: (np.float32, 3), OSP_FB_POSITION: (np.float32, 3), OSP_FB_ID_PRIMITIVE: (np.uint32, 1),
    OSP_FB_ID_OBJECT: (np.uint32, 1), OSP_FB_ID_INSTANCE: (np.uint32, 1),
}


class FrameBuffer(Object):
    kind = OSP_FRAMEBUFFER

    def __init__(self, width, height, fmt=OSP_FB_SRGBA, channels=OSP_FB_COLOR):
        super().__init__(lib().ospNewFrameBuffer(width, height, fmt, channels))
        self.width, self.height, self.format, self.channels = width, height, fmt, channels

    def reset_accumulation(self):
        lib().ospResetAccumulation(self.h)

    def render_frame(self, renderer, camera, world):
        f = lib().ospRenderFrame(self.h, renderer.h, camera.h, world.h)
        check_errors()
        return Future(f)

    def variance(self):
        return float(lib().ospGetVariance(self.h))

    def map(self, channel=OSP_FB_COLOR):
        """copy of a channel as a numpy array (H, W[, C]); row 0 is the bottom row like in OSPRay"""
        L = lib()
        p = L.ospMapFrameBuffer(self.h, channel)
        check_errors()
        if not p:
            raise OSPRayError(f"channel {channel} not available")
        n = self.width * self.height
        if channel == OSP_FB_COLOR:
            if self.format == OSP_FB_RGBA32F:
                out = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), (self.height, self.width, 4)).copy()
            else:
                out = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (self.height, self.width, 4)).copy()
        else:
            dt, comps = _CHANNEL_DTYPE[channel]
            ct = C.c_float if dt == np.float32 else C.c_uint32
            shape = (self.height, self.width) if comps == 1 else (self.height, self.width, comps)
            out = np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape).copy()
        L.ospUnmapFrameBuffer(p, self.h)
        return out

    def mapped(self, channel=OSP_FB_COLOR):
        """context manager: zero-copy numpy view of ospMapFrameBuffer's pointer, valid until the block ends
        (ospUnmapFrameBuffer) - what a C application does; map() additionally copies into a fresh array"""
        fbo = self

        class _Mapped:
            def __enter__(self_m):
                L = lib()
                self_m.p = L.ospMapFrameBuffer(fbo.h, channel)
                check_errors()
                if not self_m.p:
                    raise OSPRayError(f"channel {channel} not available")
                if channel == OSP_FB_COLOR:
                    ct, shape = (C.c_float, (fbo.height, fbo.width, 4)) if fbo.format == OSP_FB_RGBA32F else (C.c_uint8, (fbo.height, fbo.width, 4))
                else:
                    dt, comps = _CHANNEL_DTYPE[channel]
                    ct = C.c_float if dt == np.float32 else C.c_uint32
                    shape = (fbo.height, fbo.width) if comps == 1 else (fbo.height, fbo.width, comps)
                return np.ctypeslib.as_array(C.cast(self_m.p, C.POINTER(ct)), shape)

            def __exit__(self_m, *exc):
                lib().ospUnmapFrameBuffer(self_m.p, fbo.h)
                return False

        return _Mapped()

    def device_pointer(self, channel=OSP_FB_COLOR):
        bpp = C.c_uint64(0)
        p = ext().ospcudaFrameBufferDevicePointer(self.h, channel, C.byref(bpp))
        return p, int(bpp.value)

    def fused_ray_count(self):
        """rays of the last frame traced inside the pathtracer's fused tail kernel (subset of ray_counts())"""
        return int(ext().ospcudaGetFusedRayCount(self.h))

    def ray_counts(self):
        c = (C.c_uint64 * 3)()
        ext().ospcudaGetRayCounts(self.h, c)
        return int(c[0]), int(c[1]), int(c[2])


def set_render_stream(stream_handle):
    """render on an external cudaStream_t (torch.cuda.Stream().cuda_stream); None restores the own stream"""
    ext().ospcudaSetRenderStream(stream_handle)


def kernel_launch_count():
    return int(ext().ospcudaKernelLaunchCount())
