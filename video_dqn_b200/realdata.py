"""Real-data staging: `data.feather` -> label table -> pinned uint8 frame ring ->
reference-format batches for `QLearner.step` / `BatchStager.push`.

Host-side mirror of `dataloaders/q_learning_real.py:14-98` (QLearningRealDataset) and of the
`DataLoader(..., shuffle=True)` the reference wraps it in (`train_q_network.py:90-113`):

* the table written by `dataset/process_episodes_real.py:136-181` (columns `before_image`,
  `after_image`, `ep_id`, `im_start`, `im_stop`, `detector_score0-4`, `sparse_reward0-4`,
  `steps_to_reward0-4`, `steps_to_reward_neg0-4`, `inverse_actions`) is read once and every label is
  computed for the whole table up front (vectorised) instead of per `__getitem__`:
  reward = terminal = detector_score > threshold (`:15-18,79-83`), valid_mask = 1 (`:84`),
  gt = gamma ** steps_to_reward with NaN where the class is never reached (`:85-89`, value learning
  only; a NaN scalar otherwise), action = `inverse_actions` | 0 (`:90-97`);
* frames stop at uint8: JPEG decode + `Resize(224)` + `CenterCrop(224)` exactly as
  `util/torch.py:5-12` does them (same torchvision transforms on the PIL image), but NOT
  `ToTensor()/Normalize` -- `x/255`, `-mean`, `/std` are fused into the first GPU kernel
  (`vdqn_stem_pack_u8`, bit-identical) -- so a frame is 147 KB instead of 588 KB on the host, in
  pinned memory, and over PCIe;
* decoded frames land directly in a ring of pinned batches (thread pool: PIL releases the GIL while
  decoding), so `BatchStager.push` can start the asynchronous H2D without a staging copy.

`previous_images` (`:57-71`): four frames per state, ids `max(id - i, im_start)`, shape
[B, 4, 224, 224, 3].
"""
from __future__ import annotations

import os
import threading
import re
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Tuple

import numpy as np
import torch

# confidence thresholds of the five object classes (dataloaders/q_learning_real.py:15-18)
DETECTION_THRESHOLDS = np.array([0.9700177907943726, 0.9738382697105408, 0.9512060284614563,
                                 0.7334915995597839, 0.7058018445968628])
NUM_CLASSES = 5
_FRAME_RE = re.compile(r"(.*?/)(\d+).jpg")


def _multi_get(table, name: str) -> np.ndarray:
    """columns name0, name1, ... as one [rows, k] array (util/pd.py:10-14)"""
    cols = [c for c in table.columns if re.match(f"{name}\\d+$", c)]
    return np.stack([np.asarray(table[f"{name}{i}"]) for i in range(len(cols))], axis=1)


class QuadrupletTable:
    """Labels of every (s, a, r, s') row of a `data.feather`, computed once."""

    def __init__(self, location: str, *, one_action: bool = False, value_learning: bool = False,
                 inverse_actions: bool = False, previous_images: bool = False, confidence_reward: bool = False,
                 gamma: float = 0.99, root: Optional[str] = None):
        import pandas as pd
        t = pd.read_feather(location)
        self.root = os.path.dirname(os.path.abspath(location)) if root is None else root
        self.before = list(t["before_image"])
        self.after = list(t["after_image"])
        self.im_start = np.asarray(t["im_start"]) if "im_start" in t.columns else None
        self.previous_images = previous_images
        det = _multi_get(t, "detector_score")
        if confidence_reward:
            # the reference assigns `termainl` (sic) here and then returns `reward` twice (:78-80,98)
            self.reward = det
        else:
            self.reward = (det > DETECTION_THRESHOLDS).astype(np.int64)
        self.terminal = self.reward                       # `return ..., reward, reward, ...` (:98)
        self.valid_mask = np.ones_like(self.reward)
        if value_learning:
            steps = _multi_get(t, "steps_to_reward").astype(np.float64)
            gt = np.power(np.ones_like(steps) * gamma, steps)
            gt[steps == np.inf] = np.nan
            self.gt = gt
        else:
            self.gt = np.full((len(t),), np.nan)
        if inverse_actions:
            self.action = np.asarray(t["inverse_actions"]).astype(np.int64).reshape(-1)
        elif one_action:
            self.action = np.zeros(len(t), dtype=np.int64)
        else:
            raise Exception("not implemented")           # q_learning_real.py:96-97
        self.frames_per_state = 4 if previous_images else 1
        self._sparse_reward = _multi_get(t, "sparse_reward") if "sparse_reward0" in t.columns else None

    def __len__(self):
        return len(self.before)

    def reward_percentage(self) -> float:
        """fraction of rows with a sparse reward for any class (dataloaders/q_learning_real.py:51-53; printed
        as 'Reward Ratio' by the trainer, train_q_network.py:110)"""
        if self._sparse_reward is None:
            raise KeyError("data.feather has no sparse_reward0..4 columns")
        return float((self._sparse_reward.max(axis=1) > 0).sum() / self._sparse_reward.shape[0])

    def frame_paths(self, index: int) -> Tuple[List[str], List[str]]:
        """files of state s and s' of row `index` (1 or 4 each)"""
        def expand(path):
            if not self.previous_images:
                return [path]
            m = _FRAME_RE.match(path)
            prefix, im_id = m[1], int(m[2])
            start = int(self.im_start[index])
            return [prefix + "%04d.jpg" % max(im_id - i, start) for i in range(4)]
        return expand(self.before[index]), expand(self.after[index])

    def resolve(self, path: str) -> str:
        return path if os.path.isabs(path) else os.path.join(self.root, path)


_resize = None
_resize_lock = threading.Lock()


def _ensure_resize():
    """Build the Resize/CenterCrop pipeline once.  Called from the constructing (main) thread before any
    decode job is submitted: importing torchvision from several worker threads at once -- or from a
    worker while the main thread imports torchvision.models -- races inside the import system."""
    global _resize
    with _resize_lock:
        if _resize is None:
            import torchvision.transforms as T
            _resize = T.Compose([T.Resize(224), T.CenterCrop(224)])
    return _resize


def decode_frame(path: str, out: np.ndarray):
    """JPEG -> uint8 HWC [224,224,3], resized and cropped as util/torch.py:5-12 (stops before
    ToTensor/Normalize)."""
    from PIL import Image
    resize = _resize if _resize is not None else _ensure_resize()
    with Image.open(path) as im:
        out[...] = np.asarray(resize(im))


class PinnedFrameRing:
    """`depth` pinned batches; `fill(slot, rows)` decodes the frames of the given table rows into
    slot `slot` with a thread pool and returns the reference's 7-tuple
    (before, after, act, rew, term, gt, valid_mask) over that pinned memory."""

    def __init__(self, table: QuadrupletTable, batch_size: int, depth: int = 3, workers: int = 8,
                 pin: Optional[bool] = None):
        self.t, self.B, self.depth = table, batch_size, depth
        F = table.frames_per_state
        shp = (batch_size, F, 224, 224, 3) if F > 1 else (batch_size, 224, 224, 3)
        pin = torch.cuda.is_available() if pin is None else pin
        mk = lambda *s, dt: torch.empty(*s, dtype=dt, pin_memory=pin)  # noqa: E731
        gshape = table.gt.shape[1:] if table.gt.ndim > 1 else ()
        self.slots = [dict(before=mk(*shp, dt=torch.uint8), after=mk(*shp, dt=torch.uint8),
                           act=mk(batch_size, dt=torch.int64),
                           rew=mk(batch_size, NUM_CLASSES, dt=torch.from_numpy(table.reward[:1]).dtype),
                           term=mk(batch_size, NUM_CLASSES, dt=torch.from_numpy(table.reward[:1]).dtype),
                           gt=mk(batch_size, *gshape, dt=torch.float64),
                           valid=mk(batch_size, NUM_CLASSES, dt=torch.from_numpy(table.valid_mask[:1]).dtype))
                      for _ in range(depth)]
        _ensure_resize()                                  # in this thread, before the workers exist
        import PIL.Image  # noqa: F401
        self.pool = ThreadPoolExecutor(max_workers=workers)

    def fill(self, slot: int, rows) -> tuple:
        s, t = self.slots[slot % self.depth], self.t
        rows = np.asarray(rows)
        assert rows.shape[0] == self.B, "bad shape"
        jobs = []
        for which, key in ((0, "before"), (1, "after")):
            dst = s[key].numpy()
            for b, r in enumerate(rows):
                for f, path in enumerate(t.frame_paths(int(r))[which]):
                    out = dst[b, f] if t.frames_per_state > 1 else dst[b]
                    jobs.append(self.pool.submit(decode_frame, t.resolve(path), out))
        s["act"].copy_(torch.from_numpy(t.action[rows]))
        s["rew"].copy_(torch.from_numpy(t.reward[rows]))
        s["term"].copy_(torch.from_numpy(t.terminal[rows]))
        s["valid"].copy_(torch.from_numpy(t.valid_mask[rows]))
        s["gt"].copy_(torch.from_numpy(t.gt[rows]))
        for j in jobs:
            j.result()
        return s["before"], s["after"], s["act"], s["rew"], s["term"], s["gt"], s["valid"]


class QuadrupletLoader:
    """Shuffled epochs of pinned uint8 batches (drop-in for the reference's
    `DataLoader(QLearningRealDataset(...), batch_size, shuffle=True, drop_last=...)` iterator,
    train_q_network.py:100-113): `next(loader)` returns the 7-tuple of the next batch; the following
    batches are being decoded in the background (`prefetch` slots ahead)."""

    def __init__(self, table: QuadrupletTable, batch_size: int, *, seed: int = 0, prefetch: int = 2,
                 workers: int = 8, pin: Optional[bool] = None):
        if len(table) < batch_size:
            raise ValueError(f"QuadrupletLoader: the table has {len(table)} rows, fewer than one batch of {batch_size} "
                             "(incomplete batches are dropped, train_q_network.py:113)")
        self.t, self.B = table, batch_size
        # a handed-out batch stays valid until TWO further batches have been requested (its pinned
        # memory may still be the source of an asynchronous H2D copy when the next one is asked for)
        self.ring = PinnedFrameRing(table, batch_size, depth=prefetch + 2, workers=workers, pin=pin)
        self.rng = np.random.default_rng(seed)
        self.order = np.empty(0, dtype=np.int64)
        self.pos = 0
        self.issue = ThreadPoolExecutor(max_workers=1)    # fills run in issue order, decode inside is parallel
        self.pending = []
        self.slot = 0
        for _ in range(prefetch):
            self._issue()

    def _next_rows(self):
        if self.pos + self.B > len(self.order):           # new epoch (the tail that does not fill a batch is dropped)
            self.order = self.rng.permutation(len(self.t))
            self.pos = 0
        rows = self.order[self.pos:self.pos + self.B]
        self.pos += self.B
        return rows

    def _issue(self):
        rows = self._next_rows()
        self.pending.append(self.issue.submit(self.ring.fill, self.slot, rows))
        self.slot += 1

    def __iter__(self):
        return self

    def __next__(self):
        batch = self.pending.pop(0).result()
        self._issue()
        return batch
