"""Corpus embedding store: one file holding the row-major embedding matrix and the doc-id map.

The reference keeps its (sparse) index in one ``.npz`` written by ``save_sparse_csr(filename, matrix,
metadata)`` and read back by ``load_sparse_csr(filename) -> (matrix, metadata)``
(preprocessing/drqa/retriever/utils.py:21-36); ``metadata['doc_dict']`` is the ``(id -> row, row -> id)``
pair ``TfidfDocRanker`` resolves ids with (tfidf_doc_ranker.py:47,52-58).  The dense corpus of
``extract_all_emb`` (src/contrastor/utils.py:11-25: a stacked ``[rows, D]`` array) gets the same
two calls here, with a layout made for the GPU:

    [ 4096-byte header | rows x dim raw values, row-major, 4096-aligned | pickled metadata ]

* values are bf16 (2 bytes, the tcgen05 operand type) or fp32, exactly the bytes the kernels read --
  a shard is ``np.memmap``-ed and streamed to HBM through a pinned staging buffer, no parsing;
* any contiguous row range can be loaded on its own, so rank r of g reads only ``shard_bounds(N, r, g)``
   -- a 25M x 768 bf16 corpus (38.4 GB) never has to fit in host memory;
* the header carries a CRC32 of the payload; ``load_dense_index(..., verify=True)`` / ``verify(filename)``
  recompute it (a full read of the file, so it is opt-in: a rank that loads one shard of a 38 GB corpus
  should not have to read the other seven).  Truncation is always detected (sizes are checked on load).

TRUST: the metadata block is a pickle, like the reference's ``np.load(..., allow_pickle=True)``
(retriever/utils.py:33).  Unpickling runs code chosen by whoever wrote the file -- load index files only
from sources you would run code from.
"""
from __future__ import annotations

import pickle
import struct
import zlib
from typing import Optional

import numpy as np
import torch

MAGIC = b"DRSIDX01"
HEADER_BYTES = 4096
_HEADER = struct.Struct("<8sIIqqqqqI")   # magic, version, dtype, rows, dim, payload_off, payload_bytes, meta_bytes, crc32
_DTYPE_CODES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
_CODE_DTYPES = {0: (torch.float32, np.float32, 4), 1: (torch.bfloat16, np.uint16, 2), 2: (torch.float16, np.float16, 2)}


def _as_raw(t: torch.Tensor) -> np.ndarray:
    """CPU tensor -> numpy view of its bytes (bf16 has no numpy dtype: viewed as uint16)."""
    t = t.detach().cpu().contiguous()
    return t.view(torch.int16).numpy().view(np.uint16) if t.dtype == torch.bfloat16 else t.numpy()


def save_dense_index(filename: str, embeddings: torch.Tensor, metadata: Optional[dict] = None,
                     dtype: torch.dtype = torch.bfloat16, chunk_rows: int = 1 << 18) -> None:
    """utils.py:21-29 for a dense corpus.  ``embeddings`` [rows, D] (any device; converted to ``dtype``
    chunk by chunk so a device-resident corpus is never duplicated); ``metadata``: any picklable dict,
    by convention with ``'doc_dict': (id -> row, row -> id)`` like the reference."""
    if embeddings.dim() != 2:
        raise ValueError("embeddings must be [rows, dim]")
    if dtype not in _DTYPE_CODES:
        raise TypeError("dtype must be torch.bfloat16, torch.float16 or torch.float32")
    rows, dim = embeddings.shape
    esize = _CODE_DTYPES[_DTYPE_CODES[dtype]][2]
    meta = pickle.dumps(metadata, protocol=4)
    payload_bytes = rows * dim * esize
    crc = 0
    with open(filename, "wb") as f:
        f.write(b"\0" * HEADER_BYTES)
        for r0 in range(0, rows, chunk_rows):
            raw = _as_raw(embeddings[r0:r0 + chunk_rows].to(dtype)).tobytes()
            crc = zlib.crc32(raw, crc)
            f.write(raw)
        f.write(meta)
        f.seek(0)
        f.write(_HEADER.pack(MAGIC, 1, _DTYPE_CODES[dtype], rows, dim, HEADER_BYTES, payload_bytes, len(meta), crc & 0xFFFFFFFF))


def read_header(filename: str) -> dict:
    with open(filename, "rb") as f:
        head = f.read(_HEADER.size)
    if len(head) < _HEADER.size:
        raise RuntimeError(f"{filename}: not a dense index (file too short)")
    magic, version, code, rows, dim, off, nbytes, meta_bytes, crc = _HEADER.unpack(head)
    if magic != MAGIC or version != 1 or code not in _CODE_DTYPES:
        raise RuntimeError(f"{filename}: not a dense index (bad magic/version/dtype)")
    return dict(rows=rows, dim=dim, dtype=_CODE_DTYPES[code][0], payload_offset=off, payload_bytes=nbytes,
                metadata_bytes=meta_bytes, crc32=crc, _np=_CODE_DTYPES[code][1], _esize=_CODE_DTYPES[code][2])


def load_metadata(filename: str):
    h = read_header(filename)
    with open(filename, "rb") as f:
        f.seek(h["payload_offset"] + h["payload_bytes"])
        blob = f.read(h["metadata_bytes"])
    if len(blob) != h["metadata_bytes"]:
        raise RuntimeError(f"{filename}: truncated (metadata incomplete)")
    return pickle.loads(blob)


def verify(filename: str, chunk_bytes: int = 1 << 26) -> bool:
    """Recompute the payload CRC32."""
    h = read_header(filename)
    crc, left = 0, h["payload_bytes"]
    with open(filename, "rb") as f:
        f.seek(h["payload_offset"])
        while left > 0:
            buf = f.read(min(chunk_bytes, left))
            if not buf:
                return False
            crc = zlib.crc32(buf, crc)
            left -= len(buf)
    return (crc & 0xFFFFFFFF) == h["crc32"]


_verify_file = verify      # load_dense_index has a keyword of the same name


def load_rows(filename: str, lo: int = 0, hi: Optional[int] = None, device=None, chunk_rows: int = 1 << 18) -> torch.Tensor:
    """Rows [lo, hi) of the stored matrix as a tensor of the stored dtype.  ``device=None`` -> host tensor
    (I/O only; the engine itself has no CPU path); a CUDA device -> memory-mapped read streamed through a
    pinned staging buffer, ``chunk_rows`` at a time."""
    h = read_header(filename)
    hi = h["rows"] if hi is None else hi
    if not (0 <= lo <= hi <= h["rows"]):
        raise ValueError(f"row range [{lo}, {hi}) outside [0, {h['rows']})")
    dim, n = h["dim"], hi - lo
    mm = np.memmap(filename, dtype=h["_np"], mode="r", offset=h["payload_offset"] + lo * dim * h["_esize"], shape=(n, dim)) \
        if n else np.empty((0, dim), dtype=h["_np"])

    def to_tensor(a):
        t = torch.from_numpy(np.array(a))          # a private, writable copy of the mapped rows
        return t.view(torch.int16).view(torch.bfloat16) if h["dtype"] == torch.bfloat16 else t

    if device is None or torch.device(device).type == "cpu":
        return to_tensor(mm)
    dev = torch.device(device)
    out = torch.empty(n, dim, dtype=h["dtype"], device=dev)
    if n == 0:
        return out
    stage = [torch.empty(min(chunk_rows, n), dim, dtype=h["dtype"]).pin_memory() for _ in range(2)]
    events = [None, None]
    for c, r0 in enumerate(range(0, n, chunk_rows)):
        r1 = min(n, r0 + chunk_rows)
        buf = stage[c & 1]
        if events[c & 1] is not None:
            events[c & 1].synchronize()                    # the copy that last used this buffer has finished
        buf[: r1 - r0].copy_(to_tensor(mm[r0:r1]))
        out[r0:r1].copy_(buf[: r1 - r0], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        events[c & 1] = ev
    torch.cuda.current_stream(dev).synchronize()
    return out


def load_dense_index(filename: str, device=None, rank: int = 0, world_size: int = 1, verify: bool = False):
    """utils.py:32-36 for a dense corpus: ``(DenseIndex, metadata)``.  With ``world_size > 1`` the index
    holds only this rank's row shard (``shard_bounds``) with its global ``id_base``.  ``verify=True`` recomputes
    the payload CRC32 first (reads the whole file) and raises on a mismatch; the file size is always checked."""
    import os
    from .retrieval import DenseIndex, shard_bounds
    h = read_header(filename)
    expect = h["payload_offset"] + h["payload_bytes"] + h["metadata_bytes"]
    if os.path.getsize(filename) < expect:
        raise RuntimeError(f"{filename}: truncated ({os.path.getsize(filename)} bytes, header promises {expect})")
    if verify and not _verify_file(filename):
        raise RuntimeError(f"{filename}: payload CRC32 mismatch (corrupted index file)")
    lo, hi = shard_bounds(h["rows"], rank, world_size)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    emb = load_rows(filename, lo, hi, device=device)
    meta = load_metadata(filename)
    doc_ids = None
    if isinstance(meta, dict) and meta.get("doc_dict") is not None:
        doc_ids = list(meta["doc_dict"][1][lo:hi])
    return DenseIndex(emb, doc_ids, device=device, dtype=h["dtype"], id_base=lo), meta
