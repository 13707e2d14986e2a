"""Writes the layer plan of an LNet / DNet engine at one batch size as a relocatable plan file, and loads such files back
through the C ABI (``s2v_plan_*``, ``s2v_lnet_forward``, ``s2v_dnet_forward`` of include/s2v.h - csrc/plan.cu).

The file is what a host WITHOUT Python needs to run ``LNet.forward`` (reference models/LNet.py:122-139) or ``DNet.forward``
(models/DNet.py:20-28): the ordered launcher calls with their argument structs, every device pointer rewritten as
(arena, offset) - arena 0 the constant image (folded / packed weights, epilogue vectors, AdaIN pointer tables), arena 1 the
workspace (activations, statistics, I/O slots) - plus the constant image itself and the workspace regions that must be
initialised (zero padding borders, constant vectors).  The format is documented at the top of csrc/plan.cu.

    from s2v_b200 import plan_export
    plan_export.export_lnet(lnet_module, 128, "lnet_b128.s2vplan")          # once, offline (needs the GPU)
    p = plan_export.NativePlan("lnet_b128.s2vplan", device)                 # what a C host does, driven from Python
    out = p.lnet_forward(mel, face)
"""
from __future__ import annotations

import bisect
import ctypes as C
import struct

import torch

from . import _lib as L

K_I64, K_F64, K_PTR, K_BLOB, K_NULL = range(5)
ARENA_CONST, ARENA_WORK = 0, 1
_ALIGN = 256


def _tensors(obj, out):
    if isinstance(obj, torch.Tensor):
        out.append(obj)
    elif isinstance(obj, dict):
        for v in obj.values():
            _tensors(v, out)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _tensors(v, out)
    return out


class _Arenas:
    """Assigns every device storage a plan touches an offset in the constant or the workspace arena."""

    def __init__(self, engine, ent):
        self.dev = engine.dev
        self.work_keys = set()
        for t in _tensors([ent["ws"], ent["io"]], []):
            self.work_keys.add(t.untyped_storage().data_ptr())
        self.storages = {}                 # storage ptr -> (nbytes, a tensor that keeps it alive, init tag)
        self.starts = []
        for t in _tensors([ent["ws"], ent["io"], engine.W, getattr(engine, "P", {})], []):
            self._note(t)
        self.offsets = {ARENA_CONST: {}, ARENA_WORK: {}}
        self.size = {ARENA_CONST: 0, ARENA_WORK: 0}

    def _note(self, t):
        if t.device != self.dev:
            return
        st = t.untyped_storage()
        key = st.data_ptr()
        if key == 0:
            return
        tag = getattr(t, "_s2v_init", None)
        if key not in self.storages:
            self.storages[key] = [st.nbytes(), t, tag]
            bisect.insort(self.starts, key)
        elif tag is not None:
            self.storages[key][2] = tag

    def note_keep(self, keep):
        for t in _tensors(list(keep), []):
            self._note(t)

    def resolve(self, ptr: int):
        """device address -> (arena, offset); raises for an address no known tensor owns."""
        i = bisect.bisect_right(self.starts, ptr) - 1
        if i >= 0:
            key = self.starts[i]
            nbytes = self.storages[key][0]
            if key <= ptr < key + max(nbytes, 1):
                arena = ARENA_WORK if key in self.work_keys else ARENA_CONST
                offs = self.offsets[arena]
                if key not in offs:
                    offs[key] = self.size[arena]
                    self.size[arena] += -(-nbytes // _ALIGN) * _ALIGN
                return arena, offs[key] + (ptr - key)
        raise ValueError("plan export: device pointer 0x%x does not belong to the engine's weights or this plan's workspace" % ptr)

    def const_image(self) -> bytes:
        img = bytearray(self.size[ARENA_CONST])
        for key, off in self.offsets[ARENA_CONST].items():
            nbytes, t, _ = self.storages[key]
            flat = torch.empty(0, dtype=torch.uint8, device=self.dev).set_(t.untyped_storage(), 0, (nbytes,), (1,))
            img[off:off + nbytes] = flat.cpu().numpy().tobytes()
        return bytes(img)

    def inits(self):
        out = []
        for key, off in self.offsets[ARENA_WORK].items():
            nbytes, _, tag = self.storages[key]
            if tag is not None:
                out.append((off, nbytes, 0 if tag[0] == "zero" else 1, float(tag[1])))
        return out


def _struct_relocs(obj, base, arenas, out):
    """Walks a ctypes Structure: every c_void_p field becomes a relocation (and is zeroed in the blob by the caller)."""
    for name, ftype in obj._fields_:
        off = base + getattr(type(obj), name).offset
        val = getattr(obj, name)
        if isinstance(val, C.Structure):
            _struct_relocs(val, off, arenas, out)
        elif ftype is C.c_void_p:
            if val:
                arena, aoff = arenas.resolve(int(val))
                out.append((off, arena, aoff))


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _pack_relocs(relocs) -> bytes:
    return b"".join(struct.pack("<IIQ", at, arena, off) for at, arena, off in relocs)


def _encode_arg(val, ctype, arenas) -> bytes:
    if isinstance(ctype, type) and issubclass(ctype, C._Pointer):           # struct passed by address
        obj = getattr(val, "_obj", val)
        if obj is None:
            return struct.pack("<II", K_NULL, 0)
        relocs = []
        _struct_relocs(obj, 0, arenas, relocs)
        blob = bytearray(bytes(obj))
        for at, _, _ in relocs:
            blob[at:at + 8] = b"\0" * 8
        return struct.pack("<II", K_BLOB, len(blob)) + _pad8(bytes(blob)) + struct.pack("<II", len(relocs), 0) + _pack_relocs(relocs)
    if ctype is C.c_void_p:
        ptr = val.value if isinstance(val, C.c_void_p) else val
        if not ptr:
            return struct.pack("<II", K_NULL, 0)
        arena, off = arenas.resolve(int(ptr))
        return struct.pack("<II", K_PTR, 16) + struct.pack("<IIQ", 0, arena, off)
    v = getattr(val, "value", val)
    if ctype in (C.c_float, C.c_double):
        return struct.pack("<II", K_F64, 8) + struct.pack("<d", float(v))
    return struct.pack("<II", K_I64, 8) + struct.pack("<q", int(v))


def _lin_group_tables(engine, arenas):
    """Device-resident s2v_lin_group arrays hold absolute pointers: ship them as patchable tables."""
    tabs = []
    for ent in engine.W.values():
        if not (isinstance(ent, dict) and "groups" in ent):
            continue
        g = ent["groups"]
        key = g.untyped_storage().data_ptr()
        if key not in arenas.offsets[ARENA_CONST]:
            continue                                                        # not referenced by this plan
        raw = bytearray(g.cpu().numpy().tobytes())
        n = len(raw) // C.sizeof(L.LinGroup)
        arr = (L.LinGroup * n).from_buffer(raw)
        relocs = []
        for i in range(n):
            base = i * C.sizeof(L.LinGroup)
            for fld in ("wt", "bias"):
                arena, off = arenas.resolve(int(getattr(arr[i], fld)))
                relocs.append((base + getattr(L.LinGroup, fld).offset, arena, off))
        del arr
        blob = bytearray(raw)
        for at, _, _ in relocs:
            blob[at:at + 8] = b"\0" * 8
        tabs.append((arenas.offsets[ARENA_CONST][key] + g.storage_offset(), bytes(blob), relocs))
    return tabs


def export_plan(engine, ent, path: str, outputs=("out", "flow", "warp", "fake")) -> dict:
    """engine: an LNetEngine / DNetEngine; ent: one entry of its plan cache.  Returns a summary dict."""
    if ent.get("parts"):
        raise ValueError("plan export needs a single-stream plan (S2V_STREAMS=1)")
    torch.cuda.synchronize(engine.dev)
    arenas = _Arenas(engine, ent)
    ops_bin = []
    for op in ent["plan"].ops:                                              # side branch first, then the main branch: one stream
        arenas.note_keep(op.keep)
        fn = op.fn
        name = fn.__name__
        argtypes = list(fn.argtypes)[:-1]                                   # the last parameter is the stream
        if len(argtypes) != len(op.args):
            raise ValueError("%s: %d arguments for %d parameters" % (name, len(op.args), len(argtypes)))
        body = b"".join(_encode_arg(v, t, arenas) for v, t in zip(op.args, argtypes))
        ops_bin.append(struct.pack("<40sII", name.encode(), len(op.args), 0) + body)
    io_bin = []
    for name, t in ent["io"].items():
        if not isinstance(t, torch.Tensor):
            raise ValueError("plan export: I/O slot %r is not a single tensor" % name)
        assert t.is_contiguous() and t.dtype == torch.float32, name
        arena, off = arenas.resolve(t.data_ptr())
        assert arena == ARENA_WORK
        io_bin.append(struct.pack("<32sQQII", name.encode(), off, t.numel() * 4, 1 if name in outputs else 0, 0))
    tabs = _lin_group_tables(engine, arenas)          # may add the tables' targets to the constant arena: before the image
    inits = arenas.inits()
    image = arenas.const_image()
    tab_bin = [struct.pack("<QII", off, len(blob), len(rel)) + _pad8(blob) + _pack_relocs(rel) for off, blob, rel in tabs]
    head = b"S2VPLAN1" + struct.pack("<IIQQIIII", 1, len(ops_bin), len(image), arenas.size[ARENA_WORK], len(io_bin), len(inits), len(tabs), 0)
    with open(path, "wb") as f:
        f.write(head)
        f.write(b"".join(io_bin))
        f.write(b"".join(struct.pack("<QQIf", *e) for e in inits))
        f.write(b"".join(tab_bin))
        f.write(b"".join(ops_bin))
        f.write(image)
    return dict(ops=len(ops_bin), const_bytes=len(image), workspace_bytes=arenas.size[ARENA_WORK], io=list(ent["io"]))


def export_lnet(net, batch: int, path: str) -> dict:
    """net: s2v_b200.models.LNet.LNet on a CUDA device, weights loaded.  ``batch`` a multiple of 8 (or < 8)."""
    dev = next(net.parameters()).device
    net(torch.zeros(batch, 1, 80, 16, device=dev), torch.zeros(batch, 6, 96, 96, device=dev))      # builds (and warms) the plan
    eng = net.engine()
    return export_plan(eng, eng._plans[batch], path)


def export_dnet(net, batch: int, path: str, T: int = 26, stage=None) -> dict:
    dev = next(net.parameters()).device
    net(torch.zeros(batch, 3, 256, 256, device=dev), torch.zeros(batch, 73, T, device=dev), stage=stage)
    eng = net.engine()
    return export_plan(eng, eng._plans[(batch, T, "warp" if stage == "warp" else "full")], path)


class NativePlan:
    """A plan file bound to torch-owned device buffers through the C ABI only (what INTEGRATION.md's C host does)."""

    def __init__(self, path: str, device):
        self.dev = torch.device(device)
        self.lib = L.require_device(self.dev.index or 0)
        h = C.c_void_p()
        L.check(self.lib.s2v_plan_load(path.encode(), C.byref(h)), "s2v_plan_load")
        self.h = h
        with torch.cuda.device(self.dev):
            self.const = torch.empty(max(1, self.lib.s2v_plan_const_bytes(h)), dtype=torch.uint8, device=self.dev)
            self.work = torch.empty(max(1, self.lib.s2v_plan_workspace_bytes(h)), dtype=torch.uint8, device=self.dev)
            L.check(self.lib.s2v_plan_bind(h, self.const.data_ptr(), self.work.data_ptr(), self._stream()), "s2v_plan_bind")
        self.io = {}
        for i in range(self.lib.s2v_plan_num_io(h)):
            name, off, nb, out = C.c_char_p(), C.c_int64(), C.c_int64(), C.c_int()
            L.check(self.lib.s2v_plan_io_info(h, i, C.byref(name), C.byref(off), C.byref(nb), C.byref(out)), "s2v_plan_io_info")
            self.io[name.value.decode()] = (off.value, nb.value, bool(out.value))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def num_ops(self) -> int:
        return self.lib.s2v_plan_num_ops(self.h)

    def _batch(self, name, per_frame):
        return self.io[name][1] // (4 * per_frame)

    def lnet_forward(self, mel, face):
        B = self._batch("mel", 80 * 16)
        assert mel.shape == (B, 1, 80, 16) and face.shape == (B, 6, 96, 96), "this plan was exported for batch %d" % B
        mel, face = mel.float().contiguous(), face.float().contiguous()
        out = torch.empty(B, 3, 96, 96, device=self.dev)
        with torch.cuda.device(self.dev):
            L.check(self.lib.s2v_lnet_forward(self.h, mel.data_ptr(), face.data_ptr(), out.data_ptr(), self._stream()), "s2v_lnet_forward")
        return out

    def dnet_forward(self, img, coeff):
        B = self._batch("img", 3 * 256 * 256)
        assert img.shape[0] == B and coeff.shape[0] == B, "this plan was exported for batch %d" % B
        img, coeff = img.float().contiguous(), coeff.float().contiguous()
        assert coeff.numel() * 4 == self.io["coeff"][1]
        flow, warp = torch.empty(B, 2, 64, 64, device=self.dev), torch.empty(B, 3, 256, 256, device=self.dev)
        fake = torch.empty(B, 3, 256, 256, device=self.dev) if "fake" in self.io else None
        with torch.cuda.device(self.dev):
            L.check(self.lib.s2v_dnet_forward(self.h, img.data_ptr(), coeff.data_ptr(), flow.data_ptr(), warp.data_ptr(),
                                              fake.data_ptr() if fake is not None else None, self._stream()), "s2v_dnet_forward")
        out = {"flow_field": flow, "warp_image": warp}
        if fake is not None:
            out["fake_image"] = fake
        return out

    def close(self):
        if self.h:
            torch.cuda.synchronize(self.dev)
            self.lib.s2v_plan_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
