"""One binary shard format for the paired step's inputs, and its pinned-memory path to the device
(replaces the per-file .mat / CSV host stage once the features have been extracted).

Layout (little endian), every array starting on a 4096-byte boundary so that a shard can be memory-mapped, read
with O_DIRECT-sized requests, or `readinto` a page-locked buffer without an intermediate copy:

    bytes 0..7     magic  b"XMSHARD1"
    bytes 8..15    u64    length of the JSON header in bytes
    bytes 16..     JSON   {"version": 1, "rows": N, "meta": {...},
                           "arrays": {name: {"dtype": "<f4", "shape": [N, ...], "offset": o, "nbytes": b}, ...}}
    offset o       raw C-order array bytes (offsets are absolute, multiples of 4096)

All arrays of a shard share their leading dimension (`rows`: one row per paired sample / recording), so a batch is
the same row range of every array.  The reference has no such format: its loaders re-parse every file per run
(`eeg_data_utils.py:46-186`, `fmri_utils.py:115-197`); `pack_bridge_raw_dataset` converts what they return.
"""
from __future__ import annotations

import json
import os
import struct
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

MAGIC = b"XMSHARD1"
ALIGN = 4096
_DTYPES = {"<f4": np.float32, "<i8": np.int64, "<i4": np.int32, "|u1": np.uint8}

__all__ = ["write_shard", "Shard", "ShardError", "pack_bridge_raw_dataset", "host_batches"]


class ShardError(ValueError):
    pass


def _round_up(n: int, a: int = ALIGN) -> int:
    return (n + a - 1) // a * a


def write_shard(path, arrays: Dict[str, np.ndarray], meta: Optional[dict] = None) -> int:
    """Write `arrays` (same leading dimension; fp32 / int64 / int32 / uint8) to `path`; returns the file size.
    The file is written to `<path>.tmp` and renamed, so a reader never sees a partial shard."""
    if not arrays:
        raise ShardError("a shard needs at least one array")
    prepared = {}
    rows = None
    for name, a in arrays.items():
        a = np.asarray(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a)
        if a.ndim == 0:
            raise ShardError(f"{name}: arrays need a leading row dimension")
        a = np.ascontiguousarray(a)
        code = a.dtype.newbyteorder("<").str if a.dtype.itemsize > 1 else a.dtype.str
        if code not in _DTYPES:
            raise ShardError(f"{name}: dtype {a.dtype} is not supported (fp32, int64, int32, uint8)")
        if rows is None:
            rows = a.shape[0]
        elif a.shape[0] != rows:
            raise ShardError(f"{name}: {a.shape[0]} rows, other arrays have {rows}")
        prepared[name] = (a.astype(np.dtype(code), copy=False), code)  # little endian on disk
    # the header length depends on the offsets, which depend on the header length: reserve generously, then fix up
    entries = {n: {"dtype": c, "shape": list(a.shape), "offset": 0, "nbytes": int(a.nbytes)} for n, (a, c) in prepared.items()}
    header = {"version": 1, "rows": int(rows), "meta": meta or {}, "arrays": entries}
    reserve = _round_up(16 + len(json.dumps(header).encode()) + 24 * len(entries))
    off = reserve
    for n, (a, _) in prepared.items():
        entries[n]["offset"] = off
        off = _round_up(off + a.nbytes)
    blob = json.dumps(header).encode()
    assert 16 + len(blob) <= reserve
    tmp = f"{os.fspath(path)}.tmp"
    with open(tmp, "wb") as f:
        f.write(MAGIC + struct.pack("<Q", len(blob)) + blob)
        for n, (a, _) in prepared.items():
            if a.nbytes:
                f.seek(entries[n]["offset"])
                f.write(memoryview(a.reshape(-1)).cast("B"))
        f.truncate(off)
    os.replace(tmp, path)
    return off


class Shard:
    """Read side.  `shard[name]` is a read-only memory map; `read_rows` fills caller-provided (pinned) buffers
    straight from the file; `to_device` stages a row range through page-locked memory onto a CUDA device."""

    def __init__(self, path):
        self.path = os.fspath(path)
        size = os.path.getsize(self.path)
        with open(self.path, "rb") as f:
            head = f.read(16)
            if len(head) < 16 or head[:8] != MAGIC:
                raise ShardError(f"{self.path}: not an XMSHARD1 file")
            (hlen,) = struct.unpack("<Q", head[8:])
            if 16 + hlen > size:
                raise ShardError(f"{self.path}: truncated header")
            try:
                header = json.loads(f.read(hlen).decode())
            except (UnicodeDecodeError, json.JSONDecodeError) as e:
                raise ShardError(f"{self.path}: corrupt header ({e})") from None
        if header.get("version") != 1:
            raise ShardError(f"{self.path}: unsupported version {header.get('version')}")
        self.rows: int = int(header["rows"])
        self.meta: dict = header.get("meta", {})
        self.arrays: Dict[str, dict] = header["arrays"]
        for name, e in self.arrays.items():
            if e["dtype"] not in _DTYPES:
                raise ShardError(f"{self.path}: {name} has unknown dtype {e['dtype']}")
            want = int(np.prod(e["shape"], dtype=np.int64)) * np.dtype(_DTYPES[e["dtype"]]).itemsize
            if e["offset"] % ALIGN or e["nbytes"] != want or e["offset"] + e["nbytes"] > size or e["shape"][0] != self.rows:
                raise ShardError(f"{self.path}: {name} is inconsistent with the file (truncated or corrupt shard)")

    def names(self) -> List[str]:
        return list(self.arrays)

    def __len__(self) -> int:
        return self.rows

    def __getitem__(self, name: str) -> np.ndarray:
        e = self.arrays[name]
        if e["nbytes"] == 0:
            return np.empty(e["shape"], dtype=_DTYPES[e["dtype"]])
        return np.memmap(self.path, mode="r", dtype=_DTYPES[e["dtype"]], offset=e["offset"], shape=tuple(e["shape"]))

    def _row_bytes(self, name: str) -> int:
        e = self.arrays[name]
        return int(np.prod(e["shape"][1:], dtype=np.int64)) * np.dtype(_DTYPES[e["dtype"]]).itemsize

    def alloc_host(self, names: Sequence[str], rows: int, pin: Optional[bool] = None) -> Dict[str, torch.Tensor]:
        """Host buffers for `rows` rows of each array; page-locked when a CUDA device is present (pin=None)."""
        pin = torch.cuda.is_available() if pin is None else pin
        out = {}
        for n in names:
            e = self.arrays[n]
            out[n] = torch.empty((rows, *e["shape"][1:]), dtype=torch.from_numpy(np.empty(0, _DTYPES[e["dtype"]])).dtype,
                                 pin_memory=pin)
        return out

    def read_rows(self, start: int, stop: int, into: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Rows [start, stop) of every array named in `into`, read from the file directly into those buffers (their
        first stop - start rows); returns views of the filled parts."""
        if not 0 <= start <= stop <= self.rows:
            raise IndexError(f"rows [{start}, {stop}) outside [0, {self.rows})")
        n = stop - start
        out = {}
        with open(self.path, "rb", buffering=0) as f:
            for name, buf in into.items():
                e = self.arrays[name]
                if tuple(buf.shape[1:]) != tuple(e["shape"][1:]) or buf.shape[0] < n or not buf.is_contiguous():
                    raise ShardError(f"{name}: buffer {tuple(buf.shape)} does not fit rows of shape {e['shape'][1:]}")
                view = buf[:n]
                rb = self._row_bytes(name)
                if n and rb:
                    f.seek(e["offset"] + start * rb)
                    dst = memoryview(view.numpy()).cast("B")
                    got = 0
                    while got < n * rb:  # readinto may return short counts on large requests
                        k = f.readinto(dst[got:])
                        if not k:
                            raise ShardError(f"{self.path}: unexpected end of file in {name}")
                        got += k
                out[name] = view
        return out

    def to_device(self, names: Sequence[str], start: int = 0, stop: Optional[int] = None, device="cuda",
                  stream: Optional["torch.cuda.Stream"] = None) -> Dict[str, torch.Tensor]:
        """file -> page-locked host buffer -> device (asynchronous copies on `stream`, default: current stream).
        The returned tensors are ordered after the copies on that stream; the pinned staging buffers stay
        referenced by the shard until the next call, which first waits for these copies."""
        stop = self.rows if stop is None else stop
        dev = torch.device(device)
        if dev.type != "cuda":
            raise ShardError("to_device stages onto a CUDA device (no CPU path)")
        prev = getattr(self, "_staged_event", None)
        if prev is not None:
            prev.synchronize()
        host = self.read_rows(start, stop, self.alloc_host(names, stop - start, pin=True))
        stream = stream or torch.cuda.current_stream(dev)
        with torch.cuda.stream(stream):
            out = {n: h.to(dev, non_blocking=True) for n, h in host.items()}
            ev = torch.cuda.Event()
            ev.record(stream)
        self._staged, self._staged_event = host, ev
        return out


def host_batches(shard: Shard, names: Sequence[str], batch: int, drop_last: bool = True, pin: Optional[bool] = None,
                 ranks: Tuple[int, int] = (0, 1)) -> Iterator[Tuple[torch.Tensor, ...]]:
    """Consecutive `batch`-row host batches (tuples in `names` order) for `PairedTrainer.steps_from_host`, each in its
    own page-locked buffers (the trainer copies batch k+1 while batch k trains, so buffers are not recycled here).
    `ranks = (rank, world)`: rank r takes batches r, r + world, ... (batch sharding, no collective).  Every rank gets
    the SAME number of batches: the trailing `n_batches % world` batches are dropped, because each training step
    issues collectives (SyncBN, embedding exchange, gradient all-reduce) and a rank with one batch more than the
    others would wait in them forever."""
    rank, world = ranks
    if world < 1 or not 0 <= rank < world:
        raise ShardError(f"ranks=({rank}, {world}): need 0 <= rank < world")
    n_batches = shard.rows // batch if drop_last else (shard.rows + batch - 1) // batch
    n_batches -= n_batches % world
    for b in range(rank, n_batches, world):
        lo, hi = b * batch, min((b + 1) * batch, shard.rows)
        got = shard.read_rows(lo, hi, shard.alloc_host(names, hi - lo, pin=pin))
        yield tuple(got[n] for n in names)


def pack_bridge_raw_dataset(path, dataset, eeg_index: int = 0, meta: Optional[dict] = None) -> int:
    """Write a `bridge_utils.BridgeRawDataset` (or any sequence of (eeg_samples, fmri_act, fmri_conn, label, subject))
    as one shard with arrays `erp`, `pw`, `conn` (EEG sample `eeg_index` of each subject), `fmri_act`, `fmri_conn`,
    `label`, `subject`.  Every subject must have arrays of the same shapes (the reference's collate stacks them too)."""
    rows = [dataset[i] for i in range(len(dataset))]
    if not rows:
        raise ShardError("empty dataset")
    pick = [r[0][min(eeg_index, len(r[0]) - 1)] for r in rows]

    def stack(items, what):
        items = [np.asarray(x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x, dtype=np.float32) for x in items]
        if len({x.shape for x in items}) != 1:
            raise ShardError(f"{what}: subjects have different shapes {sorted({x.shape for x in items})}")
        return np.stack(items)

    arrays = {"erp": stack([p[0] for p in pick], "erp"), "pw": stack([p[1] for p in pick], "pw"),
              "conn": stack([p[2] for p in pick], "conn"), "fmri_act": stack([r[1] for r in rows], "fmri_act"),
              "fmri_conn": stack([r[2] for r in rows], "fmri_conn"),
              "label": np.asarray([int(r[3]) for r in rows], dtype=np.int64),
              "subject": np.asarray([int(r[4]) for r in rows], dtype=np.int64)}
    return write_shard(path, arrays, meta)
