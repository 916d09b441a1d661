"""Tile queue sharding across the GPUs of one box.

(tile, pair) items are independent (s2p/__init__.py:166-196): each rank takes its own items, no
data-path collective.  The only optional exchange is a gather of per-tile checksums (or rasters) to
rank 0 at the end, used by the benchmark harness and by a mosaic writer.
"""
import hashlib

import numpy as np


def shard(n_items, rank, world):
    """Indices of the items rank `rank` processes: contiguous blocks whose sizes differ by at
    most one (1521 tiles over 8 ranks -> 191,190,...)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


class DynamicQueue:
    """Dynamic sharding of a fixed list of n_items over the ranks: every rank pulls the next `chunk` indices from one
    shared counter until the list is exhausted, so a slow rank takes fewer tiles and the tail is at most one chunk
    (SURVEY.md section 8e).  The counter lives in the torch.distributed key-value store (one small round trip per
    chunk, no collective on the data path); with world == 1 it is a local integer."""

    def __init__(self, n_items, chunk, world=1, store=None, key="s2pb_tile_queue"):
        self.n, self.chunk, self.world, self.store, self.key = int(n_items), int(chunk), world, store, key
        self._local = 0
        if world > 1 and store is None:
            import torch.distributed as dist
            self.store = dist.distributed_c10d._get_default_store()

    def next(self):
        """-> list of item indices (empty when the queue is drained)"""
        if self.world == 1:
            start = self._local
            self._local += self.chunk
        else:
            start = self.store.add(self.key, self.chunk) - self.chunk
        return list(range(min(start, self.n), min(start + self.chunk, self.n)))


def checksum(a):
    """Order-independent-free 64-bit digest of a raster (NaN payloads normalised)."""
    b = np.ascontiguousarray(a)
    if b.dtype.kind == "f":
        b = np.where(np.isnan(b), np.float32(np.nan), b).astype(np.float32)
    return int.from_bytes(hashlib.blake2b(b.tobytes(), digest_size=8).digest(), "little") >> 1


def gather_checksums(local, n_items, rank, world, device=None):
    """local: {item index: checksum}.  -> on rank 0 a list of n_items checksums (None elsewhere).
    One all_gather of int64 at the end of the job; a no-op without torch.distributed."""
    out = np.full(n_items, -1, dtype=np.int64)
    for i, c in local.items():
        out[i] = c
    if world == 1:
        return out.tolist()
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(out)
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    if rank != 0:
        return None
    merged = torch.stack(parts).max(dim=0).values.cpu().numpy()
    return merged.tolist()


def process_tiles(engine, tile_ids, make_tile, dmin, dmax, params, batch=8):
    """Run the matcher over this rank's tiles in batches; make_tile(id) -> (ref, sec).
    -> {tile id: (disp checksum, valid fraction)}"""
    results = {}
    for k in range(0, len(tile_ids), batch):
        ids = tile_ids[k:k + batch]
        pairs = [make_tile(i) for i in ids]
        disp, conf, mask = engine.mgm_batch([p[0] for p in pairs], [p[1] for p in pairs], dmin, dmax, params)
        for i, d in zip(ids, disp):
            results[i] = (checksum(d), float(np.isfinite(d).mean()))
    return results
