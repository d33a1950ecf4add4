"""Minimal GGUF v3 reader / writer (numpy only).

PowerServe loads `<model>/ggml/weights.gguf` with ggml's `gguf_init_from_file(no_alloc=false)`
(/root/reference/src/model/llama/llama_model.cpp:32-36; format constants libs/ggml/include/ggml.h:265-269:
magic "GGUF", version 3, 32-byte data alignment).  The model code never reads GGUF key/values — it takes all
hyper-parameters from `model.json` (src/core/config.cpp:68-120) — so the writer emits tensors plus the two
`general.*` keys only.  The reader memory-maps the file and returns zero-copy views of the raw blocks, which is
exactly what the C-ABI's `ps_cuda_register_weight` wants (a host pointer to GGUF bytes).
"""
from __future__ import annotations

import mmap
import os
import struct
from dataclasses import dataclass
from typing import Dict, Iterable, List, Tuple

import numpy as np

GGUF_MAGIC = 0x46554747
GGUF_VERSION = 3
GGUF_ALIGNMENT = 32

# ggml type ids (libs/ggml/include/ggml.h:386-401) -> (block elements, block bytes)
GGML_F32, GGML_F16, GGML_Q4_0, GGML_Q8_0, GGML_Q4_K, GGML_Q5_K, GGML_Q6_K, GGML_Q8_K, GGML_I32 = 0, 1, 2, 8, 12, 13, 14, 15, 26
TYPE_INFO: Dict[int, Tuple[int, int]] = {
    GGML_F32: (1, 4),
    GGML_F16: (1, 2),
    GGML_Q4_0: (32, 18),
    GGML_Q8_0: (32, 34),
    GGML_Q4_K: (256, 144),
    GGML_Q5_K: (256, 176),
    GGML_Q6_K: (256, 210),
    GGML_Q8_K: (256, 292),
    GGML_I32: (1, 4),
}
TYPE_NAME = {GGML_F32: "F32", GGML_F16: "F16", GGML_Q4_0: "Q4_0", GGML_Q8_0: "Q8_0", GGML_Q4_K: "Q4_K", GGML_Q5_K: "Q5_K",
             GGML_Q6_K: "Q6_K", GGML_Q8_K: "Q8_K", GGML_I32: "I32"}

# gguf metadata value types
_T_U32, _T_STR = 4, 8


def row_bytes(ggml_type: int, ne0: int) -> int:
    blk, nbytes = TYPE_INFO[ggml_type]
    if ne0 % blk:
        raise ValueError(f"ne0={ne0} is not a multiple of the {TYPE_NAME[ggml_type]} block size {blk}")
    return ne0 // blk * nbytes


def tensor_bytes(ggml_type: int, shape: Iterable[int]) -> int:
    shape = list(shape)
    n = row_bytes(ggml_type, shape[0])
    for d in shape[1:]:
        n *= d
    return n


@dataclass
class GGUFTensor:
    name: str
    ggml_type: int
    shape: Tuple[int, ...]  # ggml order: shape[0] is the contiguous dim
    data: np.ndarray        # uint8 view of the raw bytes (zero-copy for mmap'd files)

    @property
    def nbytes(self) -> int:
        return int(self.data.nbytes)

    @property
    def host_ptr(self) -> int:
        return int(self.data.ctypes.data)


def _w_str(f, s: str) -> None:
    b = s.encode("utf-8")
    f.write(struct.pack("<Q", len(b)))
    f.write(b)


def write_gguf(path: str, tensors: List[Tuple[str, int, Tuple[int, ...], np.ndarray]], arch: str = "llama") -> None:
    """tensors: (name, ggml_type, ggml-order shape, raw bytes as a uint8/any-dtype contiguous array)."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    offs, off = [], 0
    for name, t, shape, data in tensors:
        nb = tensor_bytes(t, shape)
        raw = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        if raw.nbytes != nb:
            raise ValueError(f"{name}: got {raw.nbytes} bytes, expected {nb} for {TYPE_NAME[t]}{tuple(shape)}")
        offs.append(off)
        off += (nb + GGUF_ALIGNMENT - 1) // GGUF_ALIGNMENT * GGUF_ALIGNMENT
    with open(path, "wb") as f:
        f.write(struct.pack("<IIQQ", GGUF_MAGIC, GGUF_VERSION, len(tensors), 2))
        _w_str(f, "general.architecture"); f.write(struct.pack("<I", _T_STR)); _w_str(f, arch)
        _w_str(f, "general.alignment"); f.write(struct.pack("<II", _T_U32, GGUF_ALIGNMENT))
        for (name, t, shape, _), o in zip(tensors, offs):
            _w_str(f, name)
            f.write(struct.pack("<I", len(shape)))
            f.write(struct.pack(f"<{len(shape)}Q", *shape))
            f.write(struct.pack("<IQ", t, o))
        pad = (-f.tell()) % GGUF_ALIGNMENT
        f.write(b"\0" * pad)
        base = f.tell()
        for (name, t, shape, data), o in zip(tensors, offs):
            f.seek(base + o)
            np.ascontiguousarray(data).view(np.uint8).reshape(-1).tofile(f)
        f.truncate(base + off)  # zero-extend to the aligned end


class GGUFFile:
    """Memory-mapped GGUF v3 file; `tensors[name]` gives a zero-copy view of the raw blocks."""

    def __init__(self, path: str):
        self.path = path
        self._f = open(path, "rb")
        self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)
        buf = memoryview(self._mm)
        pos = 0

        def rd(fmt):
            nonlocal pos
            v = struct.unpack_from(fmt, buf, pos)
            pos += struct.calcsize(fmt)
            return v

        def rd_str():
            nonlocal pos
            (n,) = rd("<Q")
            s = bytes(buf[pos:pos + n]).decode("utf-8")
            pos += n
            return s

        magic, version, n_tensors, n_kv = rd("<IIQQ")
        if magic != GGUF_MAGIC:
            raise ValueError(f"{path}: bad GGUF magic {magic:#x}")
        if version not in (2, 3):
            raise ValueError(f"{path}: unsupported GGUF version {version}")
        self.kv: Dict[str, object] = {}
        scalar = {0: "<B", 1: "<b", 2: "<H", 3: "<h", 4: "<I", 5: "<i", 6: "<f", 7: "<?", 10: "<Q", 11: "<q", 12: "<d"}

        def rd_val(t):
            if t == _T_STR:
                return rd_str()
            if t == 9:  # array
                (et,) = rd("<I")
                (n,) = rd("<Q")
                return [rd_val(et) for _ in range(n)]
            return rd(scalar[t])[0]

        for _ in range(n_kv):
            k = rd_str()
            (t,) = rd("<I")
            self.kv[k] = rd_val(t)
        align = int(self.kv.get("general.alignment", GGUF_ALIGNMENT))
        infos = []
        for _ in range(n_tensors):
            name = rd_str()
            (nd,) = rd("<I")
            shape = rd(f"<{nd}Q")
            t, off = rd("<IQ")
            infos.append((name, t, tuple(int(s) for s in shape), off))
        base = (pos + align - 1) // align * align
        whole = np.frombuffer(self._mm, dtype=np.uint8)
        self.tensors: Dict[str, GGUFTensor] = {}
        for name, t, shape, off in infos:
            nb = tensor_bytes(t, shape)
            self.tensors[name] = GGUFTensor(name, t, shape, whole[base + off: base + off + nb])

    def __getitem__(self, name: str) -> GGUFTensor:
        return self.tensors[name]

    def __contains__(self, name: str) -> bool:
        return name in self.tensors
