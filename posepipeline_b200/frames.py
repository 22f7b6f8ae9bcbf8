"""Frame source for the wrappers: the reference decodes every video four times (``get_robust_reader`` test-decodes
the whole file, ``pose_pipeline/pipeline.py:62-80``, then each wrapper decodes it again synchronously on the thread that also
drives the GPU, ``wrappers/mmtrack.py:37-45`` / ``wrappers/mmpose.py:60-76``).  Here:

  * ``BlockReader``: a decode thread fills pinned host blocks and uploads them into the engine's two device slots on the
    engine's copy stream (``pe_frames_upload``) while the caller computes on the previous block -- decode, H2D and compute overlap.
    A rank of a sharded run seeks to its first frame (``CAP_PROP_POS_FRAMES``, verified, falls back to ``grab()``) instead of
    decoding everything before it.
  * ``FrameCache``: decoded frames stay resident in HBM (180 GB per B200; a 2048-frame 1080p video is 12.7 GB) keyed by a
    content fingerprint of the file, so the pose pass of a video the tracker pass has just decoded needs no decode and no H2D.
  * ``robust_reader``: ``Video.get_robust_reader`` (pipeline.py:47-87) with the same contract (fresh temp copy, every frame
    readable or ffmpeg transcode) whose validation decode is skipped when the content is already known good.

Host plumbing only: cv2 decodes on the CPU exactly as in the reference (no NVDEC requirement in north_star).
"""
from __future__ import annotations

import hashlib
import os
import queue
import threading
from collections import OrderedDict
from typing import Iterator, Optional, Tuple

import cv2
import numpy as np


def fingerprint(path: str) -> str:
    """Content key of a video file: size + SHA-1 of its first and last MiB (the temp copies every make() works on differ in
    name only)."""
    size = os.path.getsize(path)
    h = hashlib.sha1(str(size).encode())
    with open(path, "rb") as f:
        h.update(f.read(1 << 20))
        if size > (2 << 20):
            f.seek(size - (1 << 20))
            h.update(f.read(1 << 20))
    return h.hexdigest()


def pinned(shape):
    """uint8 block in page-locked host memory when torch + CUDA are present (plumbing only), else pageable."""
    try:
        import torch
        if torch.cuda.is_available():
            return torch.empty(shape, dtype=torch.uint8, pin_memory=True).numpy()
    except Exception:
        pass
    return np.empty(shape, np.uint8)


def open_at(path: str, start: int) -> cv2.VideoCapture:
    """VideoCapture positioned so that the next read() returns frame `start`."""
    cap = cv2.VideoCapture(path)
    if start > 0:
        ok = cap.set(cv2.CAP_PROP_POS_FRAMES, start) and int(round(cap.get(cv2.CAP_PROP_POS_FRAMES))) == start
        if not ok or os.environ.get("PE_FRAME_SEEK", "1") == "0":
            cap.release()
            cap = cv2.VideoCapture(path)
            for _ in range(start):                      # decode and drop (no colour conversion / copy)
                if not cap.grab():
                    break
    return cap


class Block:
    __slots__ = ("slot", "n", "first", "frames", "complete", "on_device")

    def __init__(self, slot, n, first, frames, complete, on_device):
        self.slot, self.n, self.first, self.frames, self.complete, self.on_device = slot, n, first, frames, complete, on_device


class BlockReader:
    """Iterates a video as blocks of up to `block` frames [start, stop).  A background thread decodes block k+1 into pinned
    memory and (with an engine) uploads it into device slot (k+1) % 2 while the consumer works on block k.

        for blk in reader:            # blk.frames: (n,H,W,3) uint8 host view, blk.slot: device slot (already uploading)
            reader.select(blk)        # engine stream waits for that upload
            ...compute on blk...      # synchronous engine calls
                                      # the slot is recycled when the loop asks for the next block

    `blk.complete` is False when the video ended before `stop` (the mmpose wrapper asserts on it, the mmtrack wrapper stops)."""

    def __init__(self, path: str, engine=None, block: int = 32, start: int = 0, stop: Optional[int] = None, cache_writer=None):
        self.path, self.engine, self.block, self.start, self.stop = path, engine, block, start, stop
        self.cache_writer = cache_writer        # blocks are uploaded straight into the resident cache instead of the engine's slots
        self._q: "queue.Queue" = queue.Queue()
        self._free = threading.Semaphore(2)
        self._bufs = [None, None]
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._pending: Optional[Block] = None
        self._thread.start()

    def _run(self):
        cap = None
        try:
            cap = open_at(self.path, self.start)
            i, k = self.start, 0
            while not self._stop.is_set() and (self.stop is None or i < self.stop):
                self._free.acquire()
                if self._stop.is_set():
                    break
                slot = k % 2
                want = self.block if self.stop is None else min(self.block, self.stop - i)
                n = 0
                for j in range(want):
                    ret, frame = cap.read()
                    if not ret or frame is None:
                        break
                    if self._bufs[slot] is None or self._bufs[slot].shape[1:] != frame.shape:
                        self._bufs[slot] = pinned((self.block,) + frame.shape)
                    self._bufs[slot][j] = frame
                    n += 1
                complete = n == want
                if n:
                    on_device = False
                    if self.engine is not None and hasattr(self.engine, "upload_block"):
                        dst = self.cache_writer.reserve(n) if self.cache_writer is not None else 0
                        if dst:
                            self.engine.upload_block_to(slot, dst, self._bufs[slot][:n])
                        else:
                            self.engine.upload_block(slot, self._bufs[slot][:n])
                        on_device = True
                    self._q.put(Block(slot, n, i, self._bufs[slot][:n], complete, on_device))
                else:
                    self._free.release()
                if not complete:
                    if n == 0:
                        self._q.put(Block(-1, 0, i, None, False, False))
                    break
                i += n
                k += 1
            else:
                if self.cache_writer is not None:
                    self.cache_writer.commit()              # the whole requested range was decoded
            self._q.put(None)
        except BaseException as ex:                         # surfaces in the consumer thread
            self._q.put(ex)
        finally:
            if cap is not None:
                cap.release()

    def __iter__(self) -> Iterator[Block]:
        while True:
            if self._pending is not None:                   # the previous block's slot may be reused now
                self._pending = None
                self._free.release()
            item = self._q.get()
            if item is None:
                return
            if isinstance(item, BaseException):
                raise item
            if item.slot >= 0:
                self._pending = item
            yield item
            if not item.complete:
                return

    def select(self, blk: Block):
        """Make the block the engine's staged frames (device slot if it was uploaded by the reader, else a plain staging copy)."""
        if blk.on_device:
            self.engine.select_block(blk.slot, blk.n)
        else:
            self.engine.stage_frames(blk.frames)

    def close(self):
        self._stop.set()
        self._free.release()
        self._free.release()
        self._thread.join(timeout=10)


class FrameCache:
    """Decoded frames resident in device memory, keyed by content fingerprint, LRU within a byte budget
    (``PE_FRAME_CACHE_GB``, default 24; 0 disables).  Device memory is a torch uint8 tensor (plumbing)."""

    def __init__(self, budget_bytes: Optional[int] = None):
        gb = float(os.environ.get("PE_FRAME_CACHE_GB", "24"))
        self.budget = int(gb * (1 << 30)) if budget_bytes is None else budget_bytes
        self._store: "OrderedDict[str, dict]" = OrderedDict()
        self.bytes = 0
        self.hits = self.misses = 0

    def _torch(self):
        import torch
        return torch

    def begin(self, key: str, n_frames: int, h: int, w: int, device: int = 0, first: int = 0):
        """Reserve room for frames [first, first + n_frames) of a video about to be decoded; returns a writer or None
        (disabled / does not fit)."""
        need = n_frames * h * w * 3
        if self.budget <= 0 or need > self.budget or n_frames <= 0:
            return None
        try:
            torch = self._torch()
            if not torch.cuda.is_available():
                return None
            while self.bytes + need > self.budget and self._store:
                _, old = self._store.popitem(last=False)
                self.bytes -= old["bytes"]
            buf = torch.empty((n_frames, h, w, 3), dtype=torch.uint8, device=f"cuda:{device}")
        except Exception:
            return None
        return _CacheWriter(self, key, buf, first)

    def _commit(self, key, buf, n, first=0):
        old = self._store.pop(key, None)
        if old is not None:
            self.bytes -= old["bytes"]
        self._store[key] = dict(buf=buf, n=n, first=first, bytes=buf.numel())
        self.bytes += buf.numel()

    def get(self, key: str, start: int = 0, stop: Optional[int] = None):
        """-> (device pointer of frame `start`, h, w) when frames [start, stop) of this content are resident, else None."""
        e = self._store.get(key)
        if e is None or start < e["first"] or (stop if stop is not None else e["first"] + e["n"]) > e["first"] + e["n"]:
            self.misses += 1
            return None
        self._store.move_to_end(key)
        self.hits += 1
        b = e["buf"]
        return b.data_ptr() + (start - e["first"]) * b.shape[1] * b.shape[2] * 3, b.shape[1], b.shape[2]

    def known(self, key: str) -> bool:
        return key in self._store

    def clear(self):
        self._store.clear()
        self.bytes = 0


class _CacheWriter:
    def __init__(self, cache, key, buf, first=0):
        self.cache, self.key, self.buf, self.n, self.first = cache, key, buf, 0, first
        self.frame_bytes = buf.shape[1] * buf.shape[2] * 3

    def reserve(self, n: int) -> int:
        """device address for the next n frames, or 0 when the video turned out longer than announced"""
        if self.n + n > self.buf.shape[0]:
            return 0
        p = self.buf.data_ptr() + self.n * self.frame_bytes
        self.n += n
        return p

    def commit(self):
        if self.n:
            self.cache._commit(self.key, self.buf, self.n, self.first)


CACHE = FrameCache()
_validated = set()          # fingerprints whose every frame has been decoded successfully in this process


def mark_valid(key: str):
    _validated.add(key)


def robust_reader(video_table, key, return_cap=True):
    """``Video.get_robust_reader`` (reference pipeline.py:47-87): fetch the attachment into a fresh temp .mp4, make sure every
    frame decodes (else transcode with ffmpeg), return the path or an opened capture.  The validation decode is skipped when
    this content has already been decoded completely in this process (by a previous make() or wrapper call)."""
    import shutil
    import subprocess
    import tempfile
    video = (video_table & key).fetch1("video")
    fd, outfile = tempfile.mkstemp(suffix=".mp4")
    os.close(fd)
    shutil.move(video, outfile)
    video = outfile
    fp = fingerprint(video)
    if fp not in _validated:
        cap = cv2.VideoCapture(video)
        expected_frames = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
        good = True
        for _ in range(expected_frames):
            ret, frame = cap.read()
            if not ret or frame is None:
                good = False
                break
        cap.release()
        if not good:
            fd, out2 = tempfile.mkstemp(suffix=".mp4")
            os.close(fd)
            print(f"Unable to read all the fails. Transcoding {video} to {out2}")
            subprocess.run(["ffmpeg", "-y", "-i", video, "-c:v", "libx264", "-b:v", "1M", out2])
            video = out2
        else:
            _validated.add(fp)
    if return_cap:
        cap = cv2.VideoCapture(video)
        cap.set(cv2.CAP_PROP_POS_FRAMES, 0)
        return cap
    return video
