"""weights.bin container: the model file `parakeet_create_session` loads from <model_dir>.

The reference loads three TensorRT engines from <model_dir>/{encoder,predictor,joint}.engine
(/root/reference/cpp/src/parakeet_trt.cpp:1712-1738).  This build replaces them by ONE flat
tensor container, `<model_dir>/weights.bin`, holding the NeMo state_dict tensors under their
NeMo names so a real checkpoint can later be mapped 1:1.

Layout (little endian), mirrored by csrc/weights.h:

    Header   : char magic[8]="PKB200W1"; u32 version=1; u32 n_tensors; u32 n_cfg; u32 pad
    CfgEntry : char key[32]; i64 value                                   (n_cfg times)
    TensorEnt: char name[96]; u32 dtype(0=f32,1=bf16); u32 ndim; u64 dims[4];
               u64 offset; u64 nbytes                                    (n_tensors times)
    data     : each blob 256-byte aligned, offsets absolute in the file

GEMM weight matrices are stored as bf16 (the arithmetic type of the tensor-core path);
biases, norms, depthwise kernels and position biases stay f32.
"""
from __future__ import annotations

import struct
from typing import Dict, Tuple

import numpy as np

MAGIC = b"PKB200W1"
DT_F32, DT_BF16 = 0, 1
_HDR = struct.Struct("<8sIIII")
_CFG = struct.Struct("<32sq")
_TEN = struct.Struct("<96sII4QQQ")


def f32_to_bf16_bits(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even f32 -> bf16, returned as uint16 bit patterns."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    rounded = u + (np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1)))
    return (rounded >> np.uint32(16)).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << np.uint32(16)).view(np.float32)


def write_weights(path: str, cfg: Dict[str, int], tensors: Dict[str, Tuple[np.ndarray, int]]) -> None:
    """tensors: name -> (f32 ndarray, dtype code).  bf16 tensors are rounded here."""
    names = list(tensors.keys())
    table_bytes = _HDR.size + _CFG.size * len(cfg) + _TEN.size * len(names)
    off = (table_bytes + 255) // 256 * 256
    entries, blobs = [], []
    for n in names:
        arr, dt = tensors[n]
        arr = np.ascontiguousarray(arr, dtype=np.float32)
        assert arr.ndim <= 4 and len(n.encode()) < 96, n
        blob = f32_to_bf16_bits(arr).tobytes() if dt == DT_BF16 else arr.tobytes()
        dims = list(arr.shape) + [0] * (4 - arr.ndim)
        entries.append(_TEN.pack(n.encode(), dt, arr.ndim, *dims, off, len(blob)))
        blobs.append((off, blob))
        off = (off + len(blob) + 255) // 256 * 256
    with open(path, "wb") as f:
        f.write(_HDR.pack(MAGIC, 1, len(names), len(cfg), 0))
        for k, v in cfg.items():
            assert len(k.encode()) < 32
            f.write(_CFG.pack(k.encode(), int(v)))
        for e in entries:
            f.write(e)
        for o, blob in blobs:
            f.seek(o)
            f.write(blob)
        f.truncate(off)


def read_weights(path: str):
    """Returns (cfg dict, {name: f32 ndarray}) with bf16 tensors widened to f32 exactly."""
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    magic, ver, n_t, n_c, _ = _HDR.unpack_from(mm, 0)
    if magic != MAGIC or ver != 1:
        raise ValueError(f"{path}: not a PKB200W1 container")
    pos = _HDR.size
    cfg = {}
    for _ in range(n_c):
        k, v = _CFG.unpack_from(mm, pos)
        cfg[k.rstrip(b"\0").decode()] = v
        pos += _CFG.size
    out = {}
    for _ in range(n_t):
        name, dt, nd, d0, d1, d2, d3, off, nb = _TEN.unpack_from(mm, pos)
        pos += _TEN.size
        shape = (d0, d1, d2, d3)[:nd]
        raw = mm[off:off + nb]
        if dt == DT_BF16:
            arr = bf16_bits_to_f32(np.frombuffer(raw, dtype=np.uint16)).reshape(shape)
        else:
            arr = np.frombuffer(raw, dtype=np.float32).reshape(shape).copy()
        out[name.rstrip(b"\0").decode()] = arr
    return cfg, out
