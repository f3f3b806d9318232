"""Corpus embedding store: file format round trips (host side only, no GPU)."""
import os

import numpy as np
import pytest
import torch

import drs_b200
from importlib import import_module

store = import_module(drs_b200.__name__ + ".store")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32, torch.float16])
def test_round_trip_rows_ranges_and_metadata(tmp_path, dtype):
    g = torch.Generator().manual_seed(1337)
    emb = torch.nn.functional.normalize(torch.randn(1000, 96, generator=g), dim=1)
    ids = [f"Page_{i}" for i in range(1000)]
    meta = {"doc_dict": ({d: i for i, d in enumerate(ids)}, ids), "encoder": "TFIDF-CL"}
    path = str(tmp_path / "corpus.drsidx")
    store.save_dense_index(path, emb, meta, dtype=dtype, chunk_rows=300)      # chunked write, ragged last chunk
    h = store.read_header(path)
    assert (h["rows"], h["dim"], h["dtype"]) == (1000, 96, dtype)
    assert h["payload_offset"] % 4096 == 0 and store.verify(path)
    want = emb.to(dtype)
    assert torch.equal(store.load_rows(path), want)
    assert torch.equal(store.load_rows(path, 123, 457), want[123:457])       # any row range on its own
    assert store.load_rows(path, 10, 10).shape == (0, 96)
    m = store.load_metadata(path)
    assert m["doc_dict"][1][7] == "Page_7" and m["doc_dict"][0]["Page_999"] == 999 and m["encoder"] == "TFIDF-CL"
    # the shards of a 3-rank job tile the file exactly
    parts = [store.load_rows(path, *drs_b200.shard_bounds(1000, r, 3)) for r in range(3)]
    assert torch.equal(torch.cat(parts), want)


def test_corruption_and_bad_files_are_detected(tmp_path):
    path = str(tmp_path / "c.drsidx")
    store.save_dense_index(path, torch.randn(64, 8), {"a": 1}, dtype=torch.float32)
    raw = bytearray(open(path, "rb").read())
    raw[4096 + 17] ^= 0x40
    open(path, "wb").write(bytes(raw))
    assert not store.verify(path)
    bad = str(tmp_path / "bad.bin")
    open(bad, "wb").write(b"not an index")
    with pytest.raises(RuntimeError):
        store.read_header(bad)
    with pytest.raises(ValueError):
        store.load_rows(path, 10, 100)
