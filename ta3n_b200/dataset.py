"""Feature input pipeline for the path: what sits in front of `VideoModel.forward`.

The reference (`dataset.py`) keeps one `.t7` file per frame and `torch.load`s num_segments of them per clip,
per iteration -- on real data its training loop is bound by those file opens (5 per clip), not by the model.
This module offers

* `TSNDataSet` / `VideoRecord`: drop-in for `dataset.py:16-152` (same constructor, same index rules, same
  `(num_segments*new_length, feat_dim) tensor, label` items), for code that keeps the per-frame files;
* `pack_list()` + `PackedTSNDataSet`: because `main.py:171-196` builds every split -- training included -- with
  `random_shift=False, test_mode=True`, the frames a clip contributes are a pure function of its length.  They
  are gathered ONCE into a `(num_videos, T, feat_dim)` fp32 `.npy` shard that is memory-mapped afterwards;
* `PairedFeatureLoader`: the `enumerate(zip(source_loader, target_loader))` of `main.py:343-346` with
  `RandomSampler` order, assembling every paired mini-batch directly into (pinned) staging buffers on a
  background thread, ready for `TrainStep.prefetch()`.

Host-side code only: no device work happens here.
"""
from __future__ import annotations

import json
import os
import queue
import threading
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.utils.data as data

_RGB_LIKE = ("RGB", "RGBDiff", "RGBDiff2", "RGBDiffplus")


# ---- which frames of a video are used ----------------------------------------------------------------
def _centre_ticks(num_select: int, num_segments: int) -> np.ndarray:
    # centre of each of num_segments equal spans of [0, num_select): floor(tick/2 + tick*x), same float ops
    # and order as dataset.py:98-99 / :109-110
    tick = float(num_select) / float(num_segments)
    return np.floor(tick / 2.0 + tick * np.arange(num_segments, dtype=np.float64)).astype(np.int64)


def test_segment_indices(num_frames: int, num_segments: int, new_length: int = 1) -> np.ndarray:
    """1-based start frames, `_get_test_indices` (dataset.py:103-116).  A clip shorter than
    num_segments + new_length - 1 uses its selectable frames in order and repeats the last one."""
    num_select = num_frames - new_length + 1
    if num_frames >= num_segments + new_length - 1:
        return _centre_ticks(num_select, num_segments) + 1
    if num_select <= 0:
        raise IndexError(f"video with {num_frames} frames is shorter than new_length={new_length}")   # as the reference
    return np.minimum(np.arange(num_segments, dtype=np.int64), num_select - 1) + 1


def val_segment_indices(num_frames: int, num_segments: int, new_length: int = 1) -> np.ndarray:
    """`_get_val_indices` (dataset.py:92-101): as the test rule, but a too-short clip uses frame 1 throughout."""
    if num_frames >= num_segments + new_length - 1:
        return _centre_ticks(num_frames - new_length + 1, num_segments) + 1
    return np.ones(num_segments, dtype=np.int64)


def random_segment_indices(num_frames: int, num_segments: int, new_length: int = 1) -> np.ndarray:
    """`_sample_indices` (dataset.py:77-90): one random frame per segment.  Draws from numpy's global RandomState
    with the reference's calls in the reference's order, so `numpy.random.seed` reproduces its sequences."""
    span = (num_frames - new_length + 1) // num_segments
    if span > 0:
        return np.arange(num_segments, dtype=np.int64) * span + np.random.randint(span, size=num_segments) + 1
    if num_frames > num_segments:
        return np.sort(np.random.randint(num_frames - new_length + 1, size=num_segments)).astype(np.int64) + 1
    return np.ones(num_segments, dtype=np.int64)


def expand_frames(starts: Sequence[int], num_frames: int, new_length: int = 1) -> List[int]:
    """`get` (dataset.py:128-140): new_length consecutive frames per start, not running past the last frame."""
    frames = []
    for s in starts:
        p = int(s)
        for _ in range(new_length):
            frames.append(p)
            p += 1 if p < num_frames else 0
    return frames


class VideoRecord(object):
    """One line `path num_frames label` of a list file (dataset.py:16-30)."""

    def __init__(self, row):
        self._data = row

    @property
    def path(self):
        return self._data[0]

    @property
    def num_frames(self):
        return int(self._data[1])

    @property
    def label(self):
        return int(self._data[2])


def _read_list(list_file: str, num_dataload: Optional[int]) -> List[VideoRecord]:
    with open(list_file) as f:
        records = [VideoRecord(line.strip().split(" ")) for line in f if line.strip()]
    if not records:
        raise ValueError(f"{list_file}: empty list")
    if num_dataload is None:
        return records
    # dataset.py:70-75: tile the list to exactly num_dataload items (the shorter domain is repeated, main.py:145-153)
    reps, left = divmod(int(num_dataload), len(records))
    return records * reps + records[:left]


class TSNDataSet(data.Dataset):
    """Drop-in for the reference's `TSNDataSet` (dataset.py:33-152): per-frame `.t7` feature files."""

    def __init__(self, root_path, list_file, num_dataload, num_segments=3, new_length=1, modality='RGB',
                 image_tmpl='img_{:05d}.t7', transform=None, force_grayscale=False, random_shift=True,
                 test_mode=False):
        self.root_path = root_path
        self.list_file = list_file
        self.num_segments = num_segments
        self.new_length = new_length + (1 if modality in ('RGBDiff', 'RGBDiff2', 'RGBDiffplus') else 0)   # :48-49
        self.modality = modality
        self.image_tmpl = image_tmpl
        self.transform = transform
        self.random_shift = random_shift
        self.test_mode = test_mode
        self.num_dataload = num_dataload
        self.video_list = _read_list(list_file, num_dataload)

    def segment_indices(self, record: VideoRecord) -> np.ndarray:
        if self.test_mode:
            return test_segment_indices(record.num_frames, self.num_segments, self.new_length)
        rule = random_segment_indices if self.random_shift else val_segment_indices
        return rule(record.num_frames, self.num_segments, self.new_length)

    def _load_feature(self, directory, idx):
        if self.modality in _RGB_LIKE:
            return [torch.load(os.path.join(directory, self.image_tmpl.format(idx)))]
        if self.modality == 'Flow':
            return [torch.load(os.path.join(directory, self.image_tmpl.format(axis, idx))) for axis in ('x', 'y')]
        raise ValueError(f"unknown modality {self.modality}")

    def get(self, record, indices):
        feats = []
        for p in expand_frames(indices, record.num_frames, self.new_length):
            feats.extend(self._load_feature(record.path, p))
        return torch.stack(feats), record.label

    def __getitem__(self, index):
        record = self.video_list[index]
        return self.get(record, self.segment_indices(record))

    def __len__(self):
        return len(self.video_list)


# ---- packed shards -----------------------------------------------------------------------------------
def pack_list(list_file: str, out_path: str, num_segments: int, new_length: int = 1, modality: str = 'RGB',
              image_tmpl: str = 'img_{:05d}.t7', rule: str = 'test') -> Tuple[int, int, int]:
    """Gather the frames every clip of `list_file` contributes under the deterministic index rule into one
    `(num_videos, num_segments*new_length, feat_dim)` float32 `.npy` file (+ `<out_path>.json` with labels,
    paths and the rule).  `rule`: 'test' (what main.py uses for all splits) or 'val'.  Returns the shape."""
    if rule not in ('test', 'val'):
        raise ValueError("only the deterministic rules can be packed ('test' or 'val')")
    src = TSNDataSet("", list_file, num_dataload=None, num_segments=num_segments, new_length=new_length,
                     modality=modality, image_tmpl=image_tmpl, random_shift=False, test_mode=(rule == 'test'))
    first, _ = src[0]
    shape = (len(src),) + tuple(first.shape)
    out = np.lib.format.open_memmap(out_path, mode='w+', dtype=np.float32, shape=shape)
    labels, paths, lengths = [], [], []
    for i in range(len(src)):
        x, y = (first, src.video_list[0].label) if i == 0 else src[i]
        if tuple(x.shape) != shape[1:]:
            raise ValueError(f"{src.video_list[i].path}: feature shape {tuple(x.shape)} != {shape[1:]}")
        out[i] = x.to(torch.float32).numpy()
        labels.append(int(y))
        paths.append(src.video_list[i].path)
        lengths.append(src.video_list[i].num_frames)
    out.flush()
    del out
    with open(out_path + ".json", "w") as f:
        json.dump({"list_file": os.path.abspath(list_file), "num_segments": num_segments, "new_length": src.new_length,
                   "modality": modality, "rule": rule, "labels": labels, "paths": paths, "num_frames": lengths}, f)
    return shape


class PackedTSNDataSet(data.Dataset):
    """The same items as `TSNDataSet(..., random_shift=False, test_mode=True)` served from a packed shard:
    one memory-mapped row copy instead of num_segments `torch.load` calls."""

    def __init__(self, packed_path: str, num_dataload: Optional[int] = None):
        self.features = np.load(packed_path, mmap_mode='r')
        with open(packed_path + ".json") as f:
            self.meta = json.load(f)
        if len(self.meta["labels"]) != self.features.shape[0]:
            raise ValueError("shard and metadata disagree on the number of videos")
        if self.features.dtype != np.float32 or self.features.ndim != 3:
            raise ValueError(f"{packed_path}: expected a (videos, frames, feat_dim) float32 shard as written by pack_list, "
                             f"got {self.features.dtype} {self.features.shape}")
        n = self.features.shape[0]
        # plain-ndarray view of the same mapping, one row per video: indexing a np.memmap builds a new memmap object per
        # access (3-4 us each, under the GIL), which was most of the time of a 512-row gather
        self._rows = np.asarray(self.features).reshape(n, -1)
        self.num_segments = int(self.meta["num_segments"])
        if num_dataload is None:
            self.order = np.arange(n, dtype=np.int64)
        else:
            reps, left = divmod(int(num_dataload), n)                       # dataset.py:70-75
            self.order = np.concatenate([np.tile(np.arange(n), reps), np.arange(left)]).astype(np.int64)
        self.labels = np.asarray(self.meta["labels"], dtype=np.int64)[self.order]

    def __len__(self):
        return int(self.order.shape[0])

    def __getitem__(self, index):
        row = int(self.order[index])
        return torch.from_numpy(np.array(self.features[row])), int(self.labels[index])

    def gather(self, indices: np.ndarray, out: torch.Tensor, out_labels: torch.Tensor) -> None:
        """Rows `indices` (dataset order) -> out[:len(indices)] / out_labels[:len(indices)] as ONE C-level row gather
        straight into the staging buffer (no intermediate tensors, no Python per row).  Measured on the build host, 512
        rows of 40 KB from the page cache: 3.0 ms (7 GB/s) for a Python loop of memmap row copies, 1.6 ms (13 GB/s) this
        way; more threads did not add bandwidth there."""
        rows = self.order[indices]                      # IndexError on a bad index; values are < n by construction
        k = int(rows.shape[0])
        dst = out.numpy().reshape(out.shape[0], -1)[:k]
        # mode='clip' selects numpy's unbuffered path when `out` is given ('raise' gathers into a temporary first: 3x
        # slower); nothing is ever clipped, the rows were validated by the lookup above
        np.take(self._rows, rows, axis=0, out=dst, mode='clip')
        out_labels.numpy()[:k] = self.labels[indices]


class PairedFeatureLoader:
    """`enumerate(zip(source_loader, target_loader))` of main.py:343-346 for two `PackedTSNDataSet`s: each epoch
    visits both sets in an independent random permutation (RandomSampler, main.py:188, 199), stops with the
    shorter one, and -- like DataLoader without drop_last -- ends with one short batch.  A background thread
    fills `depth` staging buffer sets (pinned when CUDA is present), so the caller only ever waits when the
    disk is slower than the GPU.

    Yields `((source_data, source_label), (target_data, target_label))`; the tensors are views of a staging
    buffer that stays untouched until TWO further batches have been requested (so an asynchronous H2D copy
    issued from it may still be in flight while the next batch is being consumed)."""

    def __init__(self, source: PackedTSNDataSet, target: PackedTSNDataSet, batch_sizes: Sequence[int],
                 seed: int = 0, depth: int = 3, pin_memory: Optional[bool] = None):
        if depth < 3:
            raise ValueError("depth must be >= 3 (current batch + previous batch + one being filled)")
        self.sets = (source, target)
        self.batch = (int(batch_sizes[0]), int(batch_sizes[1]))
        self.depth = depth
        self.gen = torch.Generator().manual_seed(seed)
        pin = torch.cuda.is_available() if pin_memory is None else bool(pin_memory)
        self.buffers = []
        for _ in range(depth):
            slot = []
            for ds, b in zip(self.sets, self.batch):
                x = torch.empty((b,) + tuple(ds.features.shape[1:]), dtype=torch.float32)
                y = torch.empty(b, dtype=torch.int64)
                slot.append((x.pin_memory(), y.pin_memory()) if pin else (x, y))
            self.buffers.append(slot)

    def __len__(self):
        return min(-(-len(ds) // b) for ds, b in zip(self.sets, self.batch))

    def __iter__(self) -> Iterator:
        perms = [torch.randperm(len(ds), generator=self.gen).numpy() for ds in self.sets]
        n_iter = len(self)
        ready: "queue.Queue" = queue.Queue()
        free = threading.Semaphore(self.depth)       # staging slots the producer may still fill
        stop = threading.Event()

        def produce():
            try:
                for it in range(n_iter):
                    while not free.acquire(timeout=0.1):
                        if stop.is_set():
                            return
                    if stop.is_set():
                        return
                    slot = self.buffers[it % self.depth]
                    sizes = []
                    for d, (ds, b) in enumerate(zip(self.sets, self.batch)):
                        idx = perms[d][it * b:(it + 1) * b]
                        ds.gather(idx, slot[d][0], slot[d][1])
                        sizes.append(len(idx))
                    ready.put((it, sizes))
                ready.put(None)
            except BaseException as e:   # surface worker failures in the consumer
                ready.put(e)

        worker = threading.Thread(target=produce, daemon=True)
        worker.start()
        try:
            while True:
                item = ready.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                it, sizes = item
                slot = self.buffers[it % self.depth]
                yield tuple((slot[d][0][:sizes[d]], slot[d][1][:sizes[d]]) for d in range(2))
                if it >= 1:
                    free.release()               # batch it-1 is no longer referenced: its slot may be refilled
        finally:
            stop.set()
            worker.join(timeout=5)


def _main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="Pack the frames a list file's clips contribute into one .npy shard")
    ap.add_argument("list_file")
    ap.add_argument("out_path")
    ap.add_argument("--num_segments", type=int, default=5)
    ap.add_argument("--new_length", type=int, default=1)
    ap.add_argument("--modality", default="RGB")
    ap.add_argument("--image_tmpl", default="img_{:05d}.t7")
    ap.add_argument("--rule", default="test", choices=["test", "val"])
    a = ap.parse_args(argv)
    shape = pack_list(a.list_file, a.out_path, a.num_segments, a.new_length, a.modality, a.image_tmpl, a.rule)
    print(f"{a.out_path}: {shape} float32, {np.prod(shape) * 4 / 1e6:.1f} MB")


if __name__ == "__main__":
    _main()
